"""Single ops of libdwb on torch CUDA tensors (thin wrappers: pointer + size marshalling only)."""
import ctypes

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _cuda(t, dtype):
    if not t.is_cuda:
        raise RuntimeError("libdwb ops need CUDA tensors (no CPU fallback)")
    return t.to(dtype).contiguous()


def cauchy_mult_sym_fwd(v, z, w):
    """Drop-in for the reference pybind op `cauchy_mult.cauchy_mult_sym_fwd(v, z, w)`
    (extensions/cauchy/cauchy.cpp:55-66): v, w (batch, N) complex64 HALF spectra, z (L) complex64,
    returns (batch, L) complex64."""
    v, z, w = _cuda(v, torch.complex64), _cuda(z, torch.complex64), _cuda(w, torch.complex64)
    if v.dim() != 2 or w.shape != v.shape or z.dim() != 1:
        raise RuntimeError("cauchy_mult_sym_fwd: v, w must be (batch, N) and z (L)")
    batch, N = v.shape
    L = z.shape[0]
    out = torch.empty(batch, L, dtype=torch.complex64, device=v.device)
    with torch.cuda.device(v.device):
        check(lib().dwb_cauchy_sym_fwd(ptr(torch.view_as_real(v)), ptr(torch.view_as_real(z)),
                                       ptr(torch.view_as_real(w)), ptr(torch.view_as_real(out)),
                                       batch, N, L, stream_ptr(v.device)))
    return out


def _c64(t):
    return ptr(torch.view_as_real(t))


def _check_vzw(name, v, z, w, dout=None):
    if v.dim() != 2 or w.shape != v.shape or z.dim() != 1 or (dout is not None and dout.shape != (v.shape[0], z.shape[0])):
        raise RuntimeError(f"{name}: v, w must be (batch, N), z (L)" + (", dout (batch, L)" if dout is not None else ""))


def cauchy_mult_fwd(v, z, w):
    """`cauchy_mult.cauchy_mult_fwd(v, z, w)` (cauchy.cpp:27-38): out[b,l] = sum_n v/(z - w), N the FULL state size."""
    v, z, w = _cuda(v, torch.complex64), _cuda(z, torch.complex64), _cuda(w, torch.complex64)
    _check_vzw("cauchy_mult_fwd", v, z, w)
    out = torch.empty(v.shape[0], z.shape[0], dtype=torch.complex64, device=v.device)
    with torch.cuda.device(v.device):
        check(lib().dwb_cauchy_fwd(_c64(v), _c64(z), _c64(w), _c64(out), v.shape[0], v.shape[1], z.shape[0], stream_ptr(v.device)))
    return out


def _bwd(fn, name, v, z, w, dout):
    v, z, w, dout = (_cuda(t, torch.complex64) for t in (v, z, w, dout))
    _check_vzw(name, v, z, w, dout)
    dv, dw = torch.empty_like(v), torch.empty_like(w)
    with torch.cuda.device(v.device):
        check(fn(_c64(v), _c64(z), _c64(w), _c64(dout), _c64(dv), _c64(dw), v.shape[0], v.shape[1], z.shape[0],
                 stream_ptr(v.device)))
    return dv, dw


def cauchy_mult_bwd(v, z, w, dout):
    """`cauchy_mult.cauchy_mult_bwd(v, z, w, dout)` -> (dv, dw)   (cauchy.cpp:40-53)"""
    return _bwd(lib().dwb_cauchy_bwd, "cauchy_mult_bwd", v, z, w, dout)


def cauchy_mult_sym_bwd(v, z, w, dout):
    """`cauchy_mult.cauchy_mult_sym_bwd(v, z, w, dout)` -> (dv, dw)   (cauchy.cpp:68-82)"""
    return _bwd(lib().dwb_cauchy_sym_bwd, "cauchy_mult_sym_bwd", v, z, w, dout)


class _CauchyMultiply(torch.autograd.Function):
    """extensions/cauchy/cauchy.py:65-86 on libdwb (no restriction on N or L)."""

    @staticmethod
    def forward(ctx, v, z, w):
        ctx.save_for_backward(v, z, w)
        return cauchy_mult_fwd(v, z, w)

    @staticmethod
    def backward(ctx, dout):
        v, z, w = ctx.saved_tensors
        dv, dw = cauchy_mult_bwd(v, z, w, dout.contiguous())
        return dv, None, dw


class _CauchyMultiplySymmetric(torch.autograd.Function):
    """extensions/cauchy/cauchy.py:89-111 on libdwb."""

    @staticmethod
    def forward(ctx, v, z, w):
        ctx.save_for_backward(v, z, w)
        return cauchy_mult_sym_fwd(v, z, w)

    @staticmethod
    def backward(ctx, dout):
        v, z, w = ctx.saved_tensors
        dv, dw = cauchy_mult_sym_bwd(v, z, w, dout.contiguous())
        return dv, None, dw


def cauchy_mult(v, z, w, symmetric=True):
    """Shape-handling, differentiable wrapper with the semantics of extensions/cauchy/cauchy.py:46-63
    (symmetric: v, w are the HALF spectra, as models/s4.py:758 passes them)."""
    v, w = torch.broadcast_tensors(v, w)
    shape = v.shape
    z = z.squeeze()
    assert z.dim() == 1
    N = v.size(-1)
    fn = _CauchyMultiplySymmetric if symmetric else _CauchyMultiply
    y = fn.apply(v.contiguous().view(-1, N), z.contiguous(), w.contiguous().view(-1, N))
    return y.view(*shape[:-1], z.size(-1))


def s4_kernel_gen(C, B, P, inv_w_real, w_imag, log_dt, l, omega=None):
    """k (2,H,l) from the stored S4 parameters (models/s4.py:674-807), evaluated in fp64 on the GPU.
    omega: complex64 (l//2+1) nodes (reference recipe) or None for exact roots of unity."""
    dev = C.device
    C, B, P = (_cuda(t, torch.float32) for t in (C, B, P))
    inv_w_real, w_imag, log_dt = (_cuda(t, torch.float32) for t in (inv_w_real, w_imag, log_dt))
    H, N = inv_w_real.shape
    om = None
    if omega is not None:
        om = torch.view_as_real(_cuda(omega, torch.complex64)).contiguous()
        assert om.shape[0] == l // 2 + 1
    k = torch.empty(2, H, l, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().dwb_s4_kernel_gen(ptr(C), ptr(B), ptr(P), ptr(inv_w_real), ptr(w_imag), ptr(log_dt), ptr(om),
                                      H, N, l, ptr(k), stream_ptr(dev)))
    return k


def fftconv_size(l):
    n = ctypes.c_int(0)
    check(lib().dwb_fftconv_size(l, ctypes.byref(n)))
    return n.value


def fftconv_prepare(k, D):
    """Cached spectrum for `fftconv` from k (2,H,l) and D (H) (or None)."""
    k = _cuda(k, torch.float32)
    _, H, l = k.shape
    n = fftconv_size(l)
    kf = torch.empty(H, n // 4 + 1, 8, dtype=torch.float32, device=k.device)
    Dd = _cuda(D.reshape(-1), torch.float32) if D is not None else None
    with torch.cuda.device(k.device):
        check(lib().dwb_fftconv_prepare(ptr(k), ptr(Dd), H, l, ptr(kf), stream_ptr(k.device)))
    return kf


def fftconv(x, kf, stats=None, part_t=None, ln_m=0.0, ln_s=1.0):
    """g = GELU(conv(y, k) + D y) with y = (ln_s*rstd)(x - mean + ln_m) + part_t  (see dwb.h)."""
    x = _cuda(x, torch.float32)
    B, H, l = x.shape
    g = torch.empty_like(x)
    psb = 0
    if part_t is not None:
        part_t = _cuda(part_t, torch.float32)
        psb = H if part_t.dim() == 2 and part_t.shape[0] == B and B > 1 else 0
    if stats is not None:
        stats = _cuda(stats, torch.float32)
        assert stats.shape == (B, l, 2)
    with torch.cuda.device(x.device):
        check(lib().dwb_fftconv(ptr(x), ptr(stats), ptr(part_t), psb, float(ln_m), float(ln_s), ptr(kf), ptr(g),
                                B, H, l, stream_ptr(x.device)))
    return g
