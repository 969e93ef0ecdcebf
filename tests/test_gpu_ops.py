"""`-m gpu`: single libdwb ops (through the C ABI) against the oracle / golden fixtures."""
import math

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


@pytest.mark.parametrize("N", [4, 8, 64, 256])
@pytest.mark.parametrize("L", [3, 17, 489, 1024, 1047])
def test_cauchy_vs_reference_complex128(dwb, N, L):
    # inputs and expected values of extensions/cauchy/test_cauchy.py:53-66 (seed 2357, batch 4);
    # the op receives the HALF spectra like models/s4.py:758
    g = load_golden("cauchy")
    v, w, z = (torch.from_numpy(g[f"{k}_{N}_{L}"]).cuda() for k in "vwz")
    out = dwb.ops.cauchy_mult_sym_fwd(v, z, w).cpu()
    ref = torch.from_numpy(g[f"out_{N}_{L}"])
    err = (out.to(torch.complex128) - ref).abs()
    assert (err.max() / ref.abs().max()).item() < 2e-5
    assert (err / ref.abs()).mean().item() < 1e-4       # the reference's own criterion is 10x KeOps + 1e-4


@pytest.mark.parametrize("N,L,batch", [(1, 5, 3), (2, 33, 2), (32, 8001, 6), (512, 3, 4), (1024, 64, 2), (32, 2 ** 16, 2),
                                       # four-outputs-per-thread kernel (batch * L >= 2^18): ragged L, odd N, N > one shared chunk
                                       (32, 16384, 32), (5, 3001, 100), (300, 1500, 180)])
def test_cauchy_shapes(dwb, N, L, batch):
    g = torch.Generator().manual_seed(N * 1000 + L)
    v = torch.randn(batch, N, dtype=torch.complex64, generator=g)
    w = torch.randn(batch, N, dtype=torch.complex64, generator=g)
    z = torch.exp(1j * torch.randn(L, generator=g)).to(torch.complex64)
    ref = O.cauchy_sym(v.cdouble(), z.cdouble(), w.cdouble())
    out = dwb.ops.cauchy_mult(v.cuda(), z.cuda(), w.cuda(), symmetric=True).cpu()
    err = (out.to(torch.complex128) - ref).abs()
    assert (err.max() / ref.abs().max()).item() < 2e-5


def test_cauchy_broadcast_wrapper(dwb):
    # the call shape of models/s4.py:758: v (2,3,H,N), w (H,N), z (L)
    g = torch.Generator().manual_seed(0)
    v = torch.randn(2, 3, 5, 32, dtype=torch.complex64, generator=g)
    w = torch.randn(5, 32, dtype=torch.complex64, generator=g) - 2
    z = torch.exp(1j * torch.randn(77, generator=g)).to(torch.complex64)
    out = dwb.ops.cauchy_mult(v.cuda(), z.cuda(), w.cuda()).cpu()
    ref = O.cauchy_sym(v.cdouble(), z.cdouble(), w.cdouble().expand(2, 3, 5, 32))
    assert out.shape == (2, 3, 5, 77)
    assert rel_max(torch.view_as_real(out), torch.view_as_real(ref)) < 2e-5


def _cauchy_f64(v, z, w, symmetric):
    """complex128 restatement of extensions/cauchy/cauchy.py:8-26 (`cauchy_mult_torch`; symmetric: half spectra)."""
    d = v.unsqueeze(-1) / (z.view(1, 1, -1) - w.unsqueeze(-1))
    if symmetric:
        d = d + v.conj().unsqueeze(-1) / (z.view(1, 1, -1) - w.conj().unsqueeze(-1))
    return d.sum(dim=-2)


@pytest.mark.parametrize("symmetric", [True, False])
@pytest.mark.parametrize("N,L,batch", [(1, 5, 3), (3, 33, 2), (32, 1000, 6), (64, 4097, 2), (256, 64, 3)])
def test_cauchy_backward_vs_autograd_complex128(dwb, symmetric, N, L, batch):
    """The four entries of the reference module `cauchy_mult` (cauchy.cpp:86-95): forward values and the (dv, dw)
    the reference's autograd.Functions must return, i.e. PyTorch autograd through the complex128 formula - the
    criterion of the reference's own extensions/cauchy/test_cauchy.py:69-99."""
    g = torch.Generator().manual_seed(17 * N + L)
    v = torch.randn(batch, N, dtype=torch.complex64, generator=g)
    w = torch.randn(batch, N, dtype=torch.complex64, generator=g) - 1.5       # poles away from the unit circle
    z = torch.exp(1j * torch.randn(L, generator=g)).to(torch.complex64)
    dout = torch.randn(batch, L, dtype=torch.complex64, generator=g)
    v64, w64 = v.cdouble().requires_grad_(), w.cdouble().requires_grad_()
    ref = _cauchy_f64(v64, z.cdouble(), w64, symmetric)
    ref.backward(dout.cdouble())
    vg, wg, zg = v.cuda().requires_grad_(), w.cuda().requires_grad_(), z.cuda()
    out = dwb.ops.cauchy_mult(vg, zg, wg, symmetric=symmetric)
    out.backward(dout.cuda())
    for name, got, want in (("out", out.detach(), ref.detach()), ("dv", vg.grad, v64.grad), ("dw", wg.grad, w64.grad)):
        e = rel_max(torch.view_as_real(got.cpu().cdouble()), torch.view_as_real(want))
        assert e < 5e-5, (name, e)
    # the shim module the reference's cauchy.py imports: same four callables, raw (batch, N) signature
    import importlib, sys
    import diffwave_sashimi_b200.shims as shims
    sys.path.insert(0, shims.PATH)
    try:
        cm = importlib.import_module("cauchy_mult")
    finally:
        sys.path.remove(shims.PATH)
    f = cm.cauchy_mult_sym_fwd if symmetric else cm.cauchy_mult_fwd
    b = cm.cauchy_mult_sym_bwd if symmetric else cm.cauchy_mult_bwd
    assert torch.equal(f(v.cuda(), zg, w.cuda()), out.detach())
    dv, dw = b(v.cuda(), zg, w.cuda(), dout.cuda())
    assert torch.equal(dv, vg.grad) and torch.equal(dw, wg.grad)


@pytest.mark.parametrize("H,L", [(4, 64), (3, 100), (2, 250), (2, 1000)])
def test_s4_kernel_gen_vs_reference(dwb, H, L):
    g = load_golden(f"s4kernel_H{H}_L{L}")
    sd = {k: v.cuda() for k, v in g["sd1"].items()}
    p = "kernel.kernel."
    args = (sd[p + "C"], sd[p + "B"], sd[p + "P"], sd[p + "inv_w_real"], sd[p + "w_imag"], sd[p + "log_dt"], L)
    k = dwb.ops.s4_kernel_gen(*args, omega=torch.from_numpy(g["omega"]).cuda()).cpu()
    assert rel_max(k, g["k"]) < 2e-5            # reference fp32 kernel, reference nodes
    sd1 = {"layer." + kk: v for kk, v in g["sd1"].items()}
    k_exact = dwb.ops.s4_kernel_gen(*args, omega=None).cpu()
    assert rel_max(k_exact, O.s4_kernel(sd1, "layer.", L, nodes="exact")) < 1e-6
    k_ref64 = O.s4_kernel(sd1, "layer.", L, nodes="reference")
    assert rel_max(k, k_ref64) < 1e-6


@pytest.mark.parametrize("L", [16000, 4000])
def test_s4_kernel_gen_full_length(dwb, L):
    H = 4
    torch.manual_seed(L)
    p = dwb.init.s4_layer_params(H)
    sd = {"layer." + k: v for k, v in p.items()}
    sd["layer.kernel.kernel.C"] = O.s4_setup_C(sd, "layer.", L).float()
    kk = "layer.kernel.kernel."
    k = dwb.ops.s4_kernel_gen(*(sd[kk + n].cuda() for n in ("C", "B", "P", "inv_w_real", "w_imag", "log_dt")), L,
                              omega=dwb.engine.reference_nodes(L).cuda()).cpu()
    ref = O.s4_kernel(sd, "layer.", L, nodes="reference")
    assert rel_max(k, ref) < 1e-6


def _fftconv_ref(x, stats, part, m, s, k, D):
    x = x.double()
    if stats is not None:
        y = (s * stats[..., 1].double())[:, None, :] * (x - stats[..., 0].double()[:, None, :] + m)
    else:
        y = x
    if part is not None:
        y = y + (part.double()[:, :, None] if part.dim() == 2 else part.double()[None, :, None])
    c = O.s4_apply(k.double(), D.double().reshape(1, -1), y)
    return torch.nn.functional.gelu(c)


@pytest.mark.parametrize("B,H,l", [(1, 1, 16), (2, 3, 40), (1, 2, 33), (3, 2, 64), (2, 2, 100), (1, 3, 250), (2, 2, 1000),
                                   (1, 2, 1001), (2, 3, 4000), (2, 2, 16000), (1, 1, 16384),
                                   # n = 32768 variants: packed split kernel (l % 4 == 0), scalar split kernel (even / odd l)
                                   (2, 2, 8200), (1, 2, 15998), (1, 2, 15999), (1, 1, 8193)])
def test_fftconv_vs_oracle(dwb, B, H, l):
    g = torch.Generator().manual_seed(B * 7 + H * 13 + l)
    x = torch.randn(B, H, l, generator=g) * 2 + 0.3
    k = torch.randn(2, H, l, generator=g) * torch.exp(-torch.arange(l) / (0.2 * l + 3))[None, None] * 0.3
    D = torch.randn(H, generator=g)
    stats = torch.stack([torch.randn(B, l, generator=g) * 0.1, torch.rand(B, l, generator=g) + 0.5], -1)
    part = torch.randn(B, H, generator=g)
    kf = dwb.ops.fftconv_prepare(k.cuda(), D.cuda())
    out = dwb.ops.fftconv(x.cuda(), kf, stats.cuda(), part.cuda(), ln_m=0.05, ln_s=1.3).cpu()
    ref = _fftconv_ref(x, stats, part, 0.05, 1.3, k, D)
    assert rel_l2(out, ref) < 2e-5 and rel_max(out, ref) < 2e-5
    # no LN / shared t-embedding row
    out2 = dwb.ops.fftconv(x.cuda(), kf, None, part[0].cuda()).cpu()
    ref2 = _fftconv_ref(x, None, part[0], 0, 1, k, D)
    assert rel_l2(out2, ref2) < 2e-5


def test_fftconv_is_linear_before_gelu_and_shift_covariant(dwb):
    # size-independent properties at the BASELINE length: conv(a x) = a conv(x) where GELU is
    # ~identity (large positive values), and an impulse reproduces the two-sided kernel itself
    l, H = 16000, 2
    g = torch.Generator().manual_seed(3)
    k = torch.randn(2, H, l, generator=g) * torch.exp(-torch.arange(l) / 500.0)[None, None]
    D = torch.zeros(H)
    kf = dwb.ops.fftconv_prepare(k.cuda(), D.cuda())
    x = torch.zeros(1, H, l)
    t0 = 7000
    x[0, :, t0] = 1.0
    bias = torch.full((H,), 0.0)
    y = dwb.ops.fftconv(x.cuda(), kf, None, bias.cuda()).cpu()
    # y = gelu(c), c[t] = k0[t - t0] (t >= t0), k1[t0 - t - 1] (t < t0)
    c = torch.zeros(H, l, dtype=torch.float64)
    c[:, t0:] = k[0, :, : l - t0].double()
    c[:, :t0] = k[1, :, :t0].flip(-1).double()
    assert rel_max(y[0], torch.nn.functional.gelu(c)) < 2e-5


def test_fftconv_rejects_too_long(dwb):
    with pytest.raises(RuntimeError):
        dwb.ops.fftconv_size(20000)


_VARIANT_SNIPPET = r'''
import sys, torch
sys.path.insert(0, {root!r})
sys.path.insert(0, {tests!r})
import diffwave_sashimi_b200 as dwb
from oracle import diffwave_oracle as O
worst = 0.0
for (B, H, l) in [(2, 2, 16000), (2, 3, 4000), (2, 2, 1000)]:
    g = torch.Generator().manual_seed(l)
    x = torch.randn(B, H, l, generator=g) * 2 + 0.3
    k = torch.randn(2, H, l, generator=g) * torch.exp(-torch.arange(l) / (0.2 * l + 3))[None, None] * 0.3
    D = torch.randn(H, generator=g)
    stats = torch.stack([torch.randn(B, l, generator=g) * 0.1, torch.rand(B, l, generator=g) + 0.5], -1)
    part = torch.randn(B, H, generator=g)
    kf = dwb.ops.fftconv_prepare(k.cuda(), D.cuda())
    out = dwb.ops.fftconv(x.cuda(), kf, stats.cuda(), part.cuda(), ln_m=0.05, ln_s=1.3).cpu().double()
    y = (1.3 * stats[..., 1].double())[:, None, :] * (x.double() - stats[..., 0].double()[:, None, :] + 0.05) + part.double()[:, :, None]
    ref = torch.nn.functional.gelu(O.s4_apply(k.double(), D.double().reshape(1, -1), y))
    worst = max(worst, ((out - ref).norm() / ref.norm()).item())
print("WORST", worst)
'''


@pytest.mark.parametrize("env", [{"DWB_FFT": "v1"}, {"DWB_FFT": "v2"}, {"DWB_FFT_WIDE": "1"}, {"DWB_FFT12": "1"},
                                 {"DWB_FFT5": "1"}])
def test_fftconv_variants(env):
    """Every S4-convolution kernel family behind the environment switches (read once per process, hence the
    subprocess): scalar one-transform (v1), scalar split (v2), packed split with the 32 B table, packed split at
    n = 8192, packed unsplit below n = 32768.  The default (packed split, 16 B table, at n = 32768; v1 below) is what
    test_fftconv_vs_oracle exercises."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = _VARIANT_SNIPPET.format(root=root, tests=os.path.join(root, "tests"))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = float([ln for ln in r.stdout.splitlines() if ln.startswith("WORST")][-1].split()[1])
    print(env, "worst rel_l2", worst)
    assert worst < 2e-5


@pytest.mark.parametrize("tag", ["lj", "small"])
def test_mel_front_end_vs_reference_golden(dwb, tag):
    """GPU mel front end (dwb_mel_spectrogram) against TacotronSTFT.mel_spectrogram of the reference's own stft.py."""
    import ast
    from diffwave_sashimi_b200 import mel as M
    g = load_golden("mel_frontend")
    kw = ast.literal_eval(str(g[f"kw_{tag}"]))
    wav, ref = torch.from_numpy(g[f"wav_{tag}"]), torch.from_numpy(g[f"mel_{tag}"])
    st = M.TacotronSTFT(**kw)
    got = st.mel_spectrogram(wav.cuda()).cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    print(f"mel front end {tag}: max abs err of log-mel {err:.2e}")
    assert err < 2e-4
    # int16-valued input path of Mel2Samp.get_mel (audio / 32768) and the oracle on the same clip
    one = M.get_mel(st, (wav[0] * 32768.0).cuda()).cpu()
    assert (one - ref[0]).abs().max().item() < 2e-4
    melb = torch.from_numpy(M.mel_filterbank(kw["sampling_rate"], kw["filter_length"], kw["n_mel_channels"], kw["mel_fmin"], kw["mel_fmax"]))
    orc = O.mel_spectrogram(wav, melb, kw["filter_length"], kw["hop_length"], kw["win_length"])
    assert (got.double() - orc).abs().max().item() < 2e-4
