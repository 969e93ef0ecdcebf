"""Parameter initialisation for fresh models (cold path; host-side torch, float64 where it matters).

Needed because `construct_model(cfg)` must hand back a usable module on a box that has neither
the reference nor a checkpoint (bench.py, smoke()).  Written from the maths in SURVEY.md
Appendix A "Initialisation"; the reference spreads the same computation over
models/s4.py:266-274 (LegS transition), :316-318 (rank-1 correction), :342-406 (normal + low rank
diagonalisation), :1150-1218 (SSKernel parameter shapes) and :1346 (D).
"""
import math

import torch


def hippo_legs_nplr(d_state: int = 64):
    """HiPPO-LegS in normal-plus-low-rank form, conjugate-pair half:
    returns w (N/2) complex64, P (N/2) complex64, B (N/2) complex64 with A = V (diag(w) - P P^*) V^*."""
    N = d_state
    q = torch.arange(N, dtype=torch.float64)
    r = torch.sqrt(2 * q + 1)
    A = -torch.tril(r[:, None] * r[None, :], -1) - torch.diag(q + 1)      # LegS transition
    B = r.clone()
    p = torch.sqrt(q + 0.5)                                                # rank-1 correction
    AP = A + p[:, None] * p[None, :]                                       # = -1/2 I + skew
    w_re = torch.diagonal(AP).mean()
    w_im, V = torch.linalg.eigh(AP.to(torch.complex128) * -1j)
    order = torch.argsort(w_im)
    w_im, V = w_im[order][: N // 2], V[:, order][:, : N // 2]
    w = torch.complex(w_re.expand(N // 2), w_im)
    Vh = V.conj().T
    Bv = Vh @ B.to(torch.complex128)
    Pv = Vh @ p.to(torch.complex128)
    return w.to(torch.complex64), Pv.to(torch.complex64), Bv.to(torch.complex64)


def s4_layer_params(H: int, d_state: int = 64, dt_min: float = 1e-3, dt_max: float = 1e-1, generator=None):
    """State-dict entries of one bidirectional S4 layer (keys relative to `<block>.layer.`)."""
    N2 = d_state // 2
    w, P, B = hippo_legs_nplr(d_state)
    g = generator
    log_dt = torch.rand(H, generator=g) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)
    C = torch.complex(torch.randn(2, H, N2, generator=g), torch.randn(2, H, N2, generator=g)) * math.sqrt(0.5)
    D = torch.randn(1, H, generator=g)
    rep = lambda v: v[None, :].expand(H, N2).contiguous()
    return {
        "D": D,
        "kernel.kernel.C": torch.view_as_real(C.contiguous()).clone(),
        "kernel.kernel.log_dt": log_dt,
        "kernel.kernel.B": torch.view_as_real(rep(B)[None].contiguous()).clone(),
        "kernel.kernel.P": torch.view_as_real(rep(P)[None].contiguous()).clone(),
        "kernel.kernel.inv_w_real": torch.log(-torch.clamp(rep(w).real, max=-1e-3)),
        "kernel.kernel.w_imag": rep(w).imag.clone(),
        "kernel.kernel.L": torch.tensor(0),
    }


def seeded_state_dict(cfg: dict, seed: int = 0, nonzero_final: bool = True):
    """Deterministic fresh weights for `cfg` as a reference-keyed state_dict (CPU tensors).
    nonzero_final: the last conv is zero-initialised in the reference (models/wavenet.py:31-36), which
    makes a fresh model output exactly 0; benchmarks and parity tests need eps to depend on the net."""
    from .models import construct_model
    torch.manual_seed(seed)
    net = construct_model(dict(cfg))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    if nonzero_final:
        g = torch.Generator().manual_seed(seed + 1)
        w = sd["final_conv.2.conv.weight"]
        sd["final_conv.2.conv.weight"] = torch.randn(w.shape, generator=g) * (1.0 / w.shape[1]) ** 0.5
        sd["final_conv.2.conv.bias"] = torch.full_like(sd["final_conv.2.conv.bias"], 0.05)
    return sd
