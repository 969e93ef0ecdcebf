"""`-m gpu`: the reverse sampler (generate.py:23-55) through the C ABI — one-step graph replay, the streaming form
behind `sampling()`, fast schedules, per-clip noise streams and the conditioning cache."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, rel_max
from oracle import diffwave_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dwb():
    import diffwave_sashimi_b200 as d
    assert torch.cuda.is_available()
    return d


def _model(dwb, cfg, sd):
    net = dwb.construct_model(dict(cfg))
    net.load_state_dict(sd)
    return net.cuda().eval()


@pytest.mark.parametrize("name", ["tiny_unet", "tiny_wnet"])
def test_streaming_steps_equal_whole_loop(dwb, name):
    """dwb_sample_steps in ragged chunks == dwb_sample, bit for bit, with and without the graph; fresh
    x_T / noise / out buffers replay the SAME captured step graph (no re-capture: launch count per call is constant)."""
    g = load_golden("traj_" + name)
    net = _model(dwb, g["cfg"], g["sd"])
    eng = net._engine_get()
    T = int(g["T"])
    dh = dwb.calc_diffusion_hyperparams(T, float(g["beta_0"]), float(g["beta_T"]), fast=True)
    coef = dwb.step_coefficients(dh)
    torch.manual_seed(int(g["seed"]))
    x_T, noise = dwb.draw_noise(tuple(g["x0"].shape), T)
    x_T, noise = x_T.cuda(), noise.cuda()
    whole = eng.sample(x_T, noise, coef)
    assert rel_l2(whole.cpu(), g["x0"]) < 1e-4
    for use_graph in (True, False):
        x = x_T.clone()
        t = T - 1
        for n in (1, 4, 2, T):                       # ragged chunk sizes, the last one clipped to what is left
            n = min(n, t + 1)
            draws = noise[T - 1 - t:].contiguous()   # draw i of the run is used at step T-1-i
            eng.sample_steps(x, draws, coef, t, n, use_graph=use_graph)
            t -= n
            if t < 0:
                break
        assert t < 0 and torch.equal(x, whole), use_graph
    l0 = eng.launch_count()
    a = eng.sample(x_T.clone(), noise.clone(), coef, out=torch.empty_like(x_T))
    l1 = eng.launch_count()
    b = eng.sample(x_T.clone(), noise.clone(), coef, out=torch.empty_like(x_T))
    l2 = eng.launch_count()
    assert torch.equal(a, whole) and torch.equal(b, whole) and l1 - l0 == l2 - l1 > 0


@pytest.mark.parametrize("name", ["tiny_unet", "tiny_wnet"])
@pytest.mark.parametrize("chunk", [1, 3, 5])
def test_sampling_chunked_staging_matches_reference(dwb, name, chunk):
    """`sampling()` with tiny staging chunks (both pinned buffers reused several times) against the reference's
    seeded generate.sampling() trajectory."""
    g = load_golden("traj_" + name)
    net = _model(dwb, g["cfg"], g["sd"])
    T = int(g["T"])
    dh = dwb.calc_diffusion_hyperparams(T, float(g["beta_0"]), float(g["beta_T"]), fast=True)
    for _ in range(2):                                # second call reuses the staging ring
        torch.manual_seed(int(g["seed"]))
        x0 = dwb.sampling(net, g["x0"].shape, dh, verbose=False, chunk_steps=chunk).cpu()
        assert rel_l2(x0, g["x0"]) < 1e-4 and rel_max(x0, g["x0"]) < 1e-4


def test_fast_schedule_through_sampler(dwb):
    """diffusion.beta list (utils.py:136-138) end to end on the GPU sampler against the oracle's loop."""
    g = load_golden("tiny_unet")
    cfg, sd = g["cfg"], g["sd"]
    net = _model(dwb, cfg, sd)
    beta = [0.0001, 0.001, 0.01, 0.05, 0.2, 0.5]
    dh = dwb.calc_diffusion_hyperparams(T=6, beta_0=1e-4, beta_T=0.02, beta=beta, fast=True)
    assert dh["T"] == 6
    B, Lx = 2, g["x"].shape[-1]
    torch.manual_seed(21)
    x0 = dwb.sampling(net, (B, 1, Lx), dh, verbose=False).cpu()
    x_T, noise = O.draw_noise(21, (B, 1, Lx), 6)
    dho = O.diffusion_schedule(6, 1e-4, 0.02, beta=beta)
    assert all(np.array_equal(dho[k].numpy(), dh[k].cpu().numpy()) for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"))
    ks = O.sashimi_kernels(cfg, sd)
    ref = O.sampling(lambda xx, tt: O.forward(cfg, sd, xx, tt, kernels=ks), x_T.double(), noise.double(), dho)
    print(f"fast schedule: rel_l2 {rel_l2(x0, ref):.2e}")
    assert rel_l2(x0, ref) < 1e-4 and rel_max(x0, ref) < 1e-4


def test_per_clip_noise_is_batch_independent(dwb):
    """A clip of a sharded/global batch == `sampling(net, (1,1,L))` under that clip's seed (what 1-GPU == N-GPU rests on)."""
    g = load_golden("traj_tiny_wnet")
    net = _model(dwb, g["cfg"], g["sd"])
    T = int(g["T"])
    dh = dwb.calc_diffusion_hyperparams(T, float(g["beta_0"]), float(g["beta_T"]), fast=True)
    Lx = g["x0"].shape[-1]
    seeds = [dwb.clip_seed(5, c) for c in range(3)]
    full = dwb.sampling(net, (3, 1, Lx), dh, verbose=False, noise=dwb.PerClipNoise(seeds)).clone()
    for c, s in enumerate(seeds):
        torch.manual_seed(s)
        one = dwb.sampling(net, (1, 1, Lx), dh, verbose=False)
        assert torch.equal(one, full[c:c + 1])
    from diffwave_sashimi_b200 import distributed as D
    again = D.generate_sharded(net, 3, Lx, dh, 5)
    assert torch.equal(again, full)


def test_cond_cache_is_not_fooled_by_recycled_storage(dwb):
    """Two same-shape mels allocated back to back usually share an address (caching allocator) and _version 0;
    the second utterance must not be vocoded with the first one's features."""
    g = load_golden("tiny_unet_cond")
    net = _model(dwb, g["cfg"], g["sd"])
    x, t = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()
    mel_np = g["mel"]
    outs, ptrs = [], []
    for scale in (1.0, -0.5):
        mel = torch.from_numpy(mel_np * scale).cuda()
        ptrs.append(mel.data_ptr())
        with torch.no_grad():
            outs.append(net((x, t), mel_spec=mel).clone())
        del mel
    print("same address reused:", ptrs[0] == ptrs[1])
    assert rel_l2(outs[0].cpu(), g["eps"]) < 1e-4
    assert not torch.allclose(outs[0], outs[1])
    sd = {k: torch.as_tensor(v) for k, v in g["sd"].items()}
    ref = O.forward(g["cfg"], sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), mel=torch.from_numpy(mel_np * -0.5))
    assert rel_l2(outs[1].cpu(), ref) < 1e-4


def test_generate_entry_point_vocodes_a_wav(dwb, tmp_path):
    """`python generate.py experiment=ljspeech ...` end to end on a synthetic 1.4 s utterance: wav -> GPU mel front end
    -> conditioning features -> T=50 sampler at audio_length = frames * hop (2.1 x the training segment: overlap-save)
    -> float32 wav files, with the reference's directory and file naming (generate.py:117-121,188-192)."""
    import os
    import subprocess
    import sys
    from scipy.io import wavfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    wavs = tmp_path / "wavs"
    wavs.mkdir()
    sr, n = 22050, 33075 - 512
    t = np.arange(n) / sr
    wav = (0.4 * np.sin(2 * np.pi * 220 * t) * np.hanning(n) * 32767).astype(np.int16)
    wavfile.write(str(wavs / "utt.wav"), sr, wav)
    r = subprocess.run([sys.executable, os.path.join(root, "generate.py"), "experiment=ljspeech", "model=sashimi_small",
                        "model.unconditional=false", "generate.random_init=true", f"dataset.data_path={wavs}",
                        "generate.mel_name=utt", "generate.n_samples=2", "generate.seed=3"],
                       capture_output=True, text=True, cwd=str(tmp_path), timeout=600,
                       env=dict(os.environ, PYTHONPATH=root))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    frames = n // 256 + 1
    assert f"begin generating audio of length {frames * 256}" in r.stdout
    outs = sorted(p for p in tmp_path.rglob("*.wav") if "waveforms" in str(p))
    assert [p.name for p in outs] == ["0k_0.wav", "0k_1.wav"], outs
    assert "_T50_betaT0.05_L16000_hop256_cond" in str(outs[0])
    for p in outs:
        rate, data = wavfile.read(str(p))
        assert rate == sr and data.dtype == np.float32 and data.shape == (frames * 256,) and np.isfinite(data).all()


def test_plan_follows_parameter_changes(dwb):
    """in-place parameter changes (optimizer step, EMA copy_, submodule load_state_dict) must not leave a
    stale compiled plan behind; modules stay deep-copyable and picklable after they have run."""
    import copy
    import pickle
    g = load_golden("tiny_wnet")
    net = _model(dwb, g["cfg"], g["sd"])
    x, t = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()
    with torch.no_grad():
        e0 = net((x, t)).clone()
        assert rel_l2(e0.cpu(), g["eps"]) < 1e-4
        net.final_conv[2].conv.weight.mul_(2.0)                  # in place, no load_state_dict
        net.final_conv[2].conv.bias.mul_(2.0)
        e1 = net((x, t)).clone()
        assert rel_l2(e1.cpu(), 2.0 * g["eps"]) < 1e-4
        sub = {k: v * 0.5 for k, v in net.final_conv[2].state_dict().items()}
        net.final_conv[2].load_state_dict(sub)                   # submodule load
        assert rel_l2(net((x, t)).cpu(), g["eps"]) < 1e-4
        twin = copy.deepcopy(net)
        assert torch.equal(twin((x, t)), net((x, t)))
        again = pickle.loads(pickle.dumps(net))
        assert torch.equal(again.cuda()((x, t)), net((x, t)))
