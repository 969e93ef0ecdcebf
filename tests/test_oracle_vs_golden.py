"""Pin oracle/diffwave_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, rel_max
from oracle import diffwave_oracle as O

TOL = 2e-5   # oracle is float64; the fixtures are the reference's fp32 outputs


@pytest.mark.parametrize("tag,kw", [
    ("T200", dict(T=200, beta_0=1e-4, beta_T=0.02)),
    ("T50", dict(T=50, beta_0=1e-4, beta_T=0.05)),
    ("fast6", dict(T=6, beta_0=1e-4, beta_T=0.02, beta=[0.0001, 0.001, 0.01, 0.05, 0.2, 0.5])),
])
def test_schedule_bit_exact(tag, kw):
    g = load_golden("schedule_" + tag)
    dh = O.diffusion_schedule(**kw)
    assert dh["T"] == int(g["T"])
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        assert np.array_equal(dh[k].numpy(), g[k]), k


def test_step_embedding():
    g = load_golden("embed")
    t = torch.from_numpy(g["t"])
    assert np.array_equal(O.step_embedding(t, 128).numpy(), g["e128"])
    assert np.array_equal(O.step_embedding(t, 16).numpy(), g["e16"])


@pytest.mark.parametrize("N", [4, 8, 64, 256])
@pytest.mark.parametrize("L", [3, 17, 489, 1024, 1047])
def test_cauchy(N, L):
    g = load_golden("cauchy")
    v, w, z = (torch.from_numpy(g[f"{k}_{N}_{L}"]) for k in "vwz")
    out = O.cauchy_sym(v.cdouble(), z.cdouble(), w.cdouble())
    ref = torch.from_numpy(g[f"out_{N}_{L}"])
    assert ((out - ref).abs() / ref.abs()).max().item() < 1e-9


@pytest.mark.parametrize("H,L", [(4, 64), (3, 100), (2, 250), (2, 1000)])
def test_s4_kernel_and_setup_C(H, L):
    g = load_golden(f"s4kernel_H{H}_L{L}")
    sd0 = {"layer." + k: v for k, v in g["sd0"].items()}
    sd1 = {"layer." + k: v for k, v in g["sd1"].items()}
    assert int(sd0["layer.kernel.kernel.L"]) == 0 and int(sd1["layer.kernel.kernel.L"]) == L
    C = O.s4_setup_C(sd0, "layer.", L)
    assert rel_max(C, sd1["layer.kernel.kernel.C"]) < 1e-4
    assert np.allclose(O.reference_nodes(L).numpy(), g["omega"], rtol=0, atol=0)
    k = O.s4_kernel(sd1, "layer.", L, nodes="reference")
    assert rel_max(k, g["k"]) < TOL
    # exact roots of unity differ from the reference's drifting complex64 nodes, but only a little
    # at these short lengths (at L=16000 the gap is 3e-3 — DESIGN.md "nodes")
    k2 = O.s4_kernel(sd1, "layer.", L, nodes="exact")
    assert rel_max(k2, g["k"]) < 5e-4


@pytest.mark.parametrize("name", ["tiny_unet", "tiny_snet", "tiny_unet_e128", "tiny_unet_cond",
                                  "tiny_unet_condB", "tiny_wnet", "tiny_wnet_cond"])
def test_tiny_forward(name):
    g = load_golden(name)
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    mel = torch.from_numpy(g["mel"]) if "mel" in g else None
    eps = O.forward(g["cfg"], g["sd"], x, t, mel)
    assert eps.shape == g["eps"].shape
    assert rel_l2(eps, g["eps"]) < TOL and rel_max(eps, g["eps"]) < TOL
    eps32 = O.forward(g["cfg"], g["sd"], x, t, mel, dtype=torch.float32)
    assert rel_l2(eps32, g["eps"]) < TOL
    if g["cfg"]["_name_"] == "sashimi":   # hoisted kernels give the same answer
        ks = O.sashimi_kernels(g["cfg"], g["sd"])
        assert rel_l2(O.forward(g["cfg"], g["sd"], x, t, mel, kernels=ks), eps) < 1e-12


@pytest.mark.parametrize("name", ["tiny_unet", "tiny_snet", "tiny_unet_cond"])
def test_setup_C_in_model(name):
    g = load_golden(name)
    lay = O.sashimi_layout(g["cfg"])
    n = 0
    for sec in "dcu":
        for (p, kind, H, l, _) in lay[sec]:
            if kind != "block":
                continue
            sd0 = dict(g["sd"])
            sd0[p + "layer.kernel.kernel.C"] = g["sd0"][p + "layer.kernel.kernel.C"]
            assert int(g["sd0"][p + "layer.kernel.kernel.L"]) == 0
            assert int(g["sd"][p + "layer.kernel.kernel.L"]) == l
            assert rel_max(O.s4_setup_C(sd0, p + "layer.", l), g["sd"][p + "layer.kernel.kernel.C"]) < 1e-4
            n += 1
    assert n > 0


@pytest.mark.parametrize("name", ["tiny_unet", "tiny_wnet"])
def test_trajectory(name):
    g = load_golden("traj_" + name)
    cfg, sd = g["cfg"], g["sd"]
    T = int(g["T"])
    dh = O.diffusion_schedule(T, float(g["beta_0"]), float(g["beta_T"]))
    x_T, noise = O.draw_noise(int(g["seed"]), g["x0"].shape, T)
    ks = O.sashimi_kernels(cfg, sd) if cfg["_name_"] == "sashimi" else None
    kw = dict(kernels=ks) if ks else {}
    x0 = O.sampling(lambda x, t: O.forward(cfg, sd, x, t, **kw), x_T.double(), noise.double(), dh)
    assert rel_l2(x0, g["x0"]) < TOL and rel_max(x0, g["x0"]) < TOL
    # and the network matters on this fixture: dropping it moves x0 by far more than the tolerance
    x0_nonet = O.sampling(lambda x, t: torch.zeros_like(x), x_T.double(), noise.double(), dh)
    assert rel_l2(x0_nonet, g["x0"]) > 1e-2


LENGTHS = [("tiny_unet", 128), ("tiny_unet", 64), ("tiny_unet", 512), ("tiny_unet", 1024), ("tiny_unet", 2304),
           ("tiny_snet", 160), ("tiny_snet", 640), ("tiny_unet_cond", 256), ("tiny_unet_cond", 1024), ("tiny_unet_cond", 2048)]


@pytest.mark.parametrize("name,Lr", LENGTHS)
def test_other_sequence_lengths(name, Lr):
    """L < l_max (kernel truncated) and L > l_max (whole kernel, FFT of size l_max + L): models/s4.py:1387,1403-1406."""
    g, gl = load_golden(name), load_golden("lengths_" + name)
    x, t = torch.from_numpy(gl[f"x_{Lr}"]), torch.from_numpy(gl["t"])
    mel = torch.from_numpy(gl[f"mel_{Lr}"]) if f"mel_{Lr}" in gl else None
    eps = O.forward(g["cfg"], g["sd"], x, t, mel)
    assert rel_l2(eps, gl[f"eps_{Lr}"]) < TOL and rel_max(eps, gl[f"eps_{Lr}"]) < TOL


@pytest.mark.parametrize("H,L", [(2, 250), (3, 64)])
def test_kernel_length_doubling(H, L):
    """models/s4.py:531-534: C <- C~ (I + dA^L) takes a kernel set up for L to 2L (and again to 4L); the host
    mirror's double_C must reproduce the reference's stored C, and the doubled kernel its k."""
    import diffwave_sashimi_b200.engine as E
    g = load_golden(f"s4double_H{H}_L{L}")
    p = "kernel.kernel."
    s1, s2, s4 = g["sd0"], g["sd1"], g["sd"]
    assert (int(s1[p + "L"]), int(s2[p + "L"]), int(s4[p + "L"])) == (L, 2 * L, 4 * L)
    args = (s1[p + "B"], s1[p + "P"], s1[p + "inv_w_real"], s1[p + "w_imag"], s1[p + "log_dt"])
    C2 = E.double_C(s1[p + "C"], *args, L)
    assert rel_max(C2, s2[p + "C"]) < 1e-4
    C4 = E.double_C(C2, *args, 2 * L)
    assert rel_max(C4, s4[p + "C"]) < 1e-4
    sd = {"layer." + k: v for k, v in s4.items()}
    sd["layer." + p + "C"] = C4
    assert rel_max(O.s4_kernel(sd, "layer.", 4 * L, nodes="reference"), g["k4"]) < 5e-5
    # rewrite_fresh_kernels: a checkpoint stored at L loaded into a model whose stage length is 4L
    blocks = [("blk.", H, 4 * L)]
    sdm = {"blk.layer." + k: v.clone() for k, v in s1.items()}
    out = list(E.rewrite_fresh_kernels(blocks, sdm))
    assert len(out) == 1 and int(sdm["blk.layer." + p + "L"]) == 4 * L and rel_max(sdm["blk.layer." + p + "C"], s4[p + "C"]) < 1e-4
    with pytest.raises(ValueError):
        list(E.rewrite_fresh_kernels([("blk.", H, 3 * L)], {"blk.layer." + k: v.clone() for k, v in s1.items()}))


@pytest.mark.parametrize("tag", ["lj", "small"])
def test_mel_front_end(tag):
    """dataloaders/stft.py TacotronSTFT.mel_spectrogram as run by the reference's own code (librosa's filterbank
    replaced by torchaudio's implementation of the same Slaney bank when the fixture was made)."""
    import ast
    from diffwave_sashimi_b200 import mel as M
    g = load_golden("mel_frontend")
    kw = ast.literal_eval(str(g[f"kw_{tag}"]))
    wav, ref = torch.from_numpy(g[f"wav_{tag}"]), g[f"mel_{tag}"]
    melb = torch.from_numpy(M.mel_filterbank(kw["sampling_rate"], kw["filter_length"], kw["n_mel_channels"], kw["mel_fmin"], kw["mel_fmax"]))
    got = O.mel_spectrogram(wav, melb, kw["filter_length"], kw["hop_length"], kw["win_length"])
    assert got.shape == ref.shape == (2, kw["n_mel_channels"], wav.shape[1] // kw["hop_length"] + 1)
    # log-mel values: absolute tolerance (the silent half sits exactly on the clamp, log 1e-5)
    assert (got - torch.from_numpy(ref).double()).abs().max().item() < 2e-4
    assert float(ref.min()) == pytest.approx(math.log(1e-5), abs=1e-6)
    # the host mirror builds the reference's float32 basis bit for bit (numpy fft of the identity x scipy window)
    step = max(1, kw["filter_length"] // 64)
    assert np.array_equal(M.forward_basis(kw["filter_length"], kw["win_length"]).numpy()[:, ::step], g[f"basis_{tag}"])
    assert (O.stft_forward_basis(kw["filter_length"], kw["win_length"]).float().numpy()[:, ::step] - g[f"basis_{tag}"]).__abs__().max() < 1e-6


def test_names():
    assert O.model_name(dict(_name_="wavenet", res_channels=256, num_res_layers=36)) == "wnet_h256_d36"
    assert O.model_name(dict(_name_="sashimi", unet=True, d_model=64, n_layers=6, pool=[4, 4], expand=2, ff=2)) \
        == "unet_d64_n6_pool_2_expand2_ff2"
