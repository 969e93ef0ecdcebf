"""Host mirror of the plan API: marshals a reference-keyed state_dict into libdwb and runs it.

All arithmetic on the per-step path happens inside libdwb.  The host side does the cold,
input-independent pieces the survey keeps in PyTorch:
  * the one-off `C` rewrite of fresh (kernel.L == 0) checkpoints    (models/s4.py:525-551)
  * the FFT node table in the reference's own complex64 recipe       (models/s4.py:553-571)
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import Config, check, lib, ptr, stream_ptr


def reference_nodes(l: int) -> torch.Tensor:
    """omega exactly as models/s4.py:561-565 builds it on the CPU: primitive root rounded to
    complex64, integer powers taken in complex64.  These drift from the true roots of unity by up
    to 1.4e-4 at l=16000 — enough to move the S4 kernels by 3e-3 — and every checkpoint was
    trained against them, so they are part of the function the engine reproduces."""
    omega = torch.tensor(np.exp(-2j * np.pi / l), dtype=torch.complex64)
    return omega ** torch.arange(0, l // 2 + 1)


def block_prefixes(cfg):
    """[(state_dict prefix, H, l)] of every DiffWaveBlock in execution order (layout of Sashimi.__init__,
    models/sashimi.py:224-275)."""
    H, l, out, i = cfg["d_model"], cfg["L"], [], 0
    for p in cfg["pool"]:
        if cfg.get("unet", True):
            for _ in range(cfg["n_layers"]):
                out.append((f"d_layers.{i}.", H, l)); i += 1
        i += 1
        l //= p; H *= cfg["expand"]
    for j in range(cfg["n_layers"]):
        out.append((f"c_layers.{j}.", H, l))
    i = 0
    for p in list(cfg["pool"])[::-1]:
        H //= cfg["expand"]; l *= p
        i += 1
        for _ in range(cfg["n_layers"]):
            out.append((f"u_layers.{i}.", H, l)); i += 1
    return out


def _dA_power(P, inv_w_real, w_imag, log_dt, L: int):
    """dA^L (H, 2N, 2N) complex128: dA the bilinear discretisation of A = diag(w~) - P~ P~^H at step dt over the
    full conjugate-pair state (models/s4.py:815-904), raised by repeated squaring (:206-224)."""
    cd = torch.complex128
    Pc = torch.view_as_complex(P.double().contiguous())[0].to(cd)         # (H,N)
    w = torch.complex(-torch.exp(inv_w_real.double()), w_imag.double())   # (H,N)
    dt = torch.exp(log_dt.double())
    H, N = w.shape
    wt, Pt = torch.cat([w, w.conj()], -1), torch.cat([Pc, Pc.conj()], -1)
    A = torch.diag_embed(wt) - Pt[:, :, None] * Pt.conj()[:, None, :]
    eye = torch.eye(2 * N, dtype=cd, device=A.device)
    s = (2.0 / dt)[:, None, None].to(cd)
    dA = torch.linalg.inv(s * eye - A) @ (s * eye + A)
    acc, base, e = eye.expand(H, -1, -1).clone(), dA, L                    # dA^L by repeated squaring
    while e:
        if e & 1:
            acc = base @ acc
        base = base @ base
        e >>= 1
    return acc


@torch.no_grad()
def setup_C(C, B, P, inv_w_real, w_imag, log_dt, L: int):
    """C <- [C~ (I - dA^L)][:N], C~ = [C, conj C] (models/s4.py:525-551).  complex128, batched over H."""
    cd = torch.complex128
    Cc = torch.view_as_complex(C.double().contiguous()).to(cd)            # (2,H,N)
    acc = _dA_power(P, inv_w_real, w_imag, log_dt, L)
    Ct = torch.cat([Cc, Cc.conj()], -1)
    Ct = Ct - torch.einsum("chn,hnm->chm", Ct, acc)
    return torch.view_as_real(Ct[..., : Cc.shape[-1]].contiguous()).float()


@torch.no_grad()
def double_C(C, B, P, inv_w_real, w_imag, log_dt, L: int):
    """Kernel-length doubling of a checkpoint whose kernels were set up for length L (models/s4.py:531-534,
    `double_length`): C <- C~ (I + dA^L), the inverse bookkeeping of `setup_C`, so that the stored parameter again
    means C~ (I - dA^{2L}) of the original C~.  complex128, batched over H."""
    cd = torch.complex128
    Cc = torch.view_as_complex(C.double().contiguous()).to(cd)
    dA_L = _dA_power(P, inv_w_real, w_imag, log_dt, L)
    Ct = torch.cat([Cc, Cc.conj()], -1)
    Ct = Ct + torch.einsum("chn,hnm->chm", Ct, dA_L)
    return torch.view_as_real(Ct[..., : Cc.shape[-1]].contiguous()).float()


@torch.no_grad()
def rewrite_fresh_kernels(blocks, sd):
    """In `sd`: bring every S4 kernel to its stage length l.  kernel.L == 0 (fresh model): the one-off C rewrite
    the reference does on its first forward (models/s4.py:525-551); kernel.L = l / 2^k (checkpoint trained on
    shorter segments): k doublings (s4.py:531-534).  Runs on whatever device the tensors live on.
    Yields (key prefix, new C, l) for every rewritten block."""
    for (p, H, l) in blocks:
        k = p + "layer.kernel.kernel."
        Lcur = int(sd[k + "L"])
        if Lcur == l:
            continue
        args = (sd[k + "B"], sd[k + "P"], sd[k + "inv_w_real"], sd[k + "w_imag"], sd[k + "log_dt"])
        if Lcur == 0:
            newC = setup_C(sd[k + "C"], *args, l)
        else:
            if Lcur > l or l % Lcur or (l // Lcur) & (l // Lcur - 1):
                raise ValueError(f"{k}L = {Lcur} cannot be doubled to the stage length {l} (models/s4.py:531-534 doubles; "
                                 f"a kernel set up for a LONGER length cannot be shortened exactly)")
            newC = sd[k + "C"]
            while Lcur < l:
                newC = double_C(newC, *args, Lcur)
                Lcur *= 2
        newC = newC.to(sd[k + "C"].device)
        sd[k + "C"] = newC
        sd[k + "L"] = torch.tensor(l)
        yield k, newC, l


def config_struct(cfg: dict) -> Config:
    """The reference's constructor kwargs (configs/model/*.yaml) as the C ABI's dwb_config."""
    sashimi = cfg["_name_"] == "sashimi"
    c = Config()
    c.model = _lib.MODEL_SASHIMI if sashimi else _lib.MODEL_WAVENET
    c.unconditional = int(bool(cfg.get("unconditional", False)))
    c.embed_in = cfg.get("diffusion_step_embed_dim_in", 128)
    c.embed_mid = cfg.get("diffusion_step_embed_dim_mid", 512)
    c.embed_out = cfg.get("diffusion_step_embed_dim_out", 512)
    c.mel_bands = 80
    if sashimi:
        pool = list(cfg["pool"])
        if len(pool) > _lib.DWB_MAX_POOL:
            raise ValueError("too many pool stages")
        c.d_model, c.n_layers, c.n_pool = cfg["d_model"], cfg["n_layers"], len(pool)
        for i, p in enumerate(pool):
            c.pool[i] = p
        c.expand, c.ff, c.unet, c.L = cfg["expand"], cfg["ff"], int(bool(cfg.get("unet", True))), cfg["L"]
        c.d_state_half = 32
    else:
        c.res_channels, c.skip_channels = cfg["res_channels"], cfg["skip_channels"]
        c.num_res_layers, c.dilation_cycle = cfg["num_res_layers"], cfg["dilation_cycle"]
    return c


class Engine:
    """One libdwb plan for one model on one device."""

    def __init__(self, cfg: dict, module_or_sd, device=None, nodes="reference"):
        self._plan = ctypes.c_void_p()
        sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else dict(module_or_sd)
        if device is None:
            device = next((v.device for v in sd.values() if v.is_cuda), None)
            if device is None:
                if not torch.cuda.is_available():
                    raise RuntimeError("diffwave_sashimi_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
                device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.cfg = dict(cfg)
        self.sashimi = cfg["_name_"] == "sashimi"
        c = config_struct(cfg)
        self._c = c
        with torch.cuda.device(self.device):
            check(lib().dwb_plan_create(ctypes.byref(c), self.device.index or 0, ctypes.byref(self._plan)))
            try:
                self._load(sd, module_or_sd if hasattr(module_or_sd, "state_dict") else None, nodes)
                check(lib().dwb_plan_finalize(self._plan, stream_ptr(self.device)))
            except Exception:
                self.close()
                raise
        self._cond_cache = None

    # ---- weights ------------------------------------------------------------------------
    def _block_prefixes(self):
        return block_prefixes(self.cfg)

    @torch.no_grad()
    def _load(self, sd, module, nodes):
        st = stream_ptr(self.device)
        if self.sashimi:
            for k, newC, l in rewrite_fresh_kernels(self._block_prefixes(), sd):
                if module is not None:   # keep the module's state identical to the reference's after a forward
                    dict(module.named_parameters())[k + "C"].copy_(newC)
                    dict(module.named_buffers())[k + "L"].fill_(l)
            if nodes != "exact":
                for l in sorted({l for (_, _, l) in self._block_prefixes()}):
                    om = reference_nodes(l) if nodes == "reference" else (
                        torch.tensor(np.exp(-2j * np.pi / l), dtype=torch.complex64, device=self.device)
                        ** torch.arange(0, l // 2 + 1, device=self.device)).cpu()
                    self._set(f"nodes.{l}", torch.view_as_real(om).contiguous(), st)
        for name, t in sd.items():
            self._set(name, t, st)

    def _set(self, name, t, st):
        t = t.detach()
        if t.dtype in (torch.int64, torch.int32, torch.int16, torch.uint8, torch.bool):
            t, dt = t.to(torch.int64), _lib.I64
        else:
            t, dt = t.to(torch.float32), _lib.F32
        t = t.contiguous()
        shape = (ctypes.c_int64 * max(1, t.dim()))(*t.shape)
        check(lib().dwb_plan_set_tensor(self._plan, name.encode(), ptr(t), dt, shape, t.dim(), int(t.is_cuda), st))

    # ---- conditioning -------------------------------------------------------------------
    def cond_layout(self, L):
        n = ctypes.c_int(0)
        check(lib().dwb_plan_cond_layout(self._plan, L, ctypes.byref(n), None, None, None))
        ch, ln, off = (ctypes.c_int * n.value)(), (ctypes.c_int * n.value)(), (ctypes.c_int64 * n.value)()
        check(lib().dwb_plan_cond_layout(self._plan, L, ctypes.byref(n), ch, ln, off))
        return list(ch), list(ln), list(off)

    @torch.no_grad()
    def cond_features(self, mel, L):
        """(cond_batch, sum_i H_i l_i) conditioning features from a mel (cb, 80, frames): per block
        two weight-normed ConvTranspose2d + leaky-ReLU(0.4), crop to the FIRST l_i samples, 1x1
        80 -> H_i (models/sashimi.py:160-175, models/wavenet.py:98-111).  t-independent."""
        # the cache entry keeps the source tensor alive and is matched by identity: a freed mel's address (and
        # _version 0) is routinely handed to the next same-shape allocation, so pointer keys would alias utterances
        c = self._cond_cache
        if c is not None and c["mel"] is mel and c["version"] == mel._version and c["L"] == L:
            return c["out"]
        src = mel
        mel = mel.to(self.device, torch.float32).contiguous()
        cb, bands, frames = mel.shape
        if bands != self._c.mel_bands:
            raise ValueError(f"mel_spec has {bands} bands, the model expects {self._c.mel_bands}")
        ch, ln, off = self.cond_layout(L)
        total = off[-1] + ch[-1] * ln[-1]
        out = torch.empty(cb * total, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().dwb_plan_cond_features(self._plan, ptr(mel), cb, frames, L, ptr(out), stream_ptr(self.device)))
        self._cond_cache = {"mel": src, "version": src._version, "L": L, "out": out}
        return out

    # ---- hot path -----------------------------------------------------------------------
    def _cond(self, mel, L):
        if mel is None:
            if not self._c.unconditional:
                raise RuntimeError("conditional model called without mel_spec")
            return None, 0
        if self._c.unconditional:
            raise RuntimeError("mel_spec passed to an unconditional model (models/sashimi.py:161 asserts the same)")
        return self.cond_features(mel, L), mel.shape[0]

    @torch.no_grad()
    def forward(self, audio, diffusion_steps, mel_spec=None):
        """eps = net((audio, diffusion_steps), mel_spec)   (generate.py:51)"""
        x = audio.to(self.device, torch.float32).contiguous()
        B, ch, L = x.shape
        assert ch == 1
        t = diffusion_steps.to(self.device, torch.float32).reshape(-1).contiguous()
        if t.numel() != B:
            raise ValueError("diffusion_steps must have one entry per batch element")
        cond, cb = self._cond(mel_spec, L)
        eps = torch.empty_like(x)
        with torch.cuda.device(self.device):
            check(lib().dwb_forward(self._plan, ptr(x), ptr(t), ptr(cond), cb, ptr(eps), B, L, stream_ptr(self.device)))
        return eps

    @torch.no_grad()
    def sample(self, x_T, noise, coef, mel_spec=None, out=None, use_graph=True):
        """x_0 of the T-step reverse loop (generate.py:47-54) from pre-drawn noise.
        x_T (B,1,L), noise (T-1,B,1,L) on the device; coef (3,T) fp32 host tensor (see dwb.h)."""
        B, _, L = x_T.shape
        T = coef.shape[1]
        assert x_T.is_cuda and x_T.is_contiguous() and x_T.dtype == torch.float32
        if T > 1:
            assert noise.is_cuda and noise.is_contiguous() and noise.dtype == torch.float32 and noise.shape[0] == T - 1
        cond, cb = self._cond(mel_spec, L)
        if out is None:
            out = torch.empty_like(x_T)
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        cp = coef.numpy().ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        with torch.cuda.device(self.device):
            check(lib().dwb_sample(self._plan, ptr(x_T), ptr(noise) if T > 1 else None, ptr(cond), cb, cp, T, ptr(out),
                                   B, L, int(use_graph), stream_ptr(self.device)))
        return out

    @torch.no_grad()
    def sample_steps(self, x, noise, coef, t_start, n_steps, mel_spec=None, use_graph=True):
        """Advance x (B,1,L), in place, by the reverse steps t_start, t_start-1, ... (n_steps of them) of the
        schedule `coef` (3,T).  noise (n_draws,B,1,L) on the device: draw i is used at step t_start - i and none
        at t = 0.  The streaming form of `sample` (dwb_sample_steps): `sampling()` draws the noise of later
        steps on the CPU generator while earlier steps run."""
        B, _, L = x.shape
        T = coef.shape[1]
        assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
        need = n_steps if t_start - n_steps + 1 > 0 else n_steps - 1
        if need > 0:
            assert noise.is_cuda and noise.is_contiguous() and noise.dtype == torch.float32 and noise.shape[0] >= need
        cond, cb = self._cond(mel_spec, L)
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        cp = coef.numpy().ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        with torch.cuda.device(self.device):
            check(lib().dwb_sample_steps(self._plan, ptr(x), ptr(noise) if need > 0 else None, ptr(cond), cb, cp, T,
                                         t_start, n_steps, B, L, int(use_graph), stream_ptr(self.device)))
        return x

    def staging(self, B, L, steps):
        """Reusable staging for `sampling()`: two pinned host chunks of `steps` noise draws, their device twins, a
        pinned x_T, a copy stream and the events that order their reuse.  Kept per (B, L, steps) so repeated calls
        neither page-fault nor pin 800 MB again."""
        key = (B, L, steps)
        st = getattr(self, "_staging", None)
        if st is None or st["key"] != key:
            with torch.cuda.device(self.device):
                st = {"key": key,
                      "host": [torch.empty((steps, B, 1, L), pin_memory=True) for _ in range(2)],
                      "dev": [torch.empty((steps, B, 1, L), device=self.device) for _ in range(2)],
                      "x_host": torch.empty((B, 1, L), pin_memory=True),
                      "copied": [None, None],       # event: H2D of host[i] finished -> host[i] may be redrawn
                      "consumed": [None, None],     # event: the steps reading dev[i] finished -> dev[i] may be overwritten
                      "x_copied": None,
                      "stream": torch.cuda.Stream(device=self.device)}
            self._staging = st
        return st

    @torch.no_grad()
    def profile(self, audio, diffusion_steps, mel_spec=None, iters=3):
        """{category: (ms per forward, launches per forward)} from CUDA events around every launch."""
        x = audio.to(self.device, torch.float32).contiguous()
        B, _, L = x.shape
        t = diffusion_steps.to(self.device, torch.float32).reshape(-1).contiguous()
        cond, cb = self._cond(mel_spec, L)
        eps = torch.empty_like(x)
        n = len(_lib.PROF_CATEGORIES)
        ms, cnt = (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
        with torch.cuda.device(self.device):
            check(lib().dwb_plan_profile(self._plan, ptr(x), ptr(t), ptr(cond), cb, ptr(eps), B, L, iters, ms, cnt,
                                         stream_ptr(self.device)))
        return {c: (ms[i] / iters, cnt[i] // iters) for i, c in enumerate(_lib.PROF_CATEGORIES) if cnt[i]}

    # ---- introspection ------------------------------------------------------------------
    def mix_block(self, block, g, x, skip=None, exact=False):
        """Channel-mixing half of DiffWaveBlock `block` on caller tensors (B,H,l): returns (out, stats).
        exact=True forces the fp32 SIMT kernel (the parity reference of the tensor-core paths)."""
        g = g.to(self.device, torch.float32).contiguous()
        x = x.to(self.device, torch.float32).contiguous()
        sk = None if skip is None else skip.to(self.device, torch.float32).contiguous()
        B, H, l = x.shape
        out = torch.empty_like(x)
        stats = torch.empty(B, l, 2, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().dwb_plan_mix_block(self._plan, block, 1 if exact else 0, ptr(g), ptr(x), ptr(sk), ptr(out),
                                           ptr(stats), B, stream_ptr(self.device)))
        return out, stats

    def launch_count(self):
        n = ctypes.c_int64(0)
        check(lib().dwb_plan_launch_count(self._plan, ctypes.byref(n)))
        return n.value

    def work(self, L):
        b, f = ctypes.c_double(0), ctypes.c_double(0)
        check(lib().dwb_plan_work(self._plan, L, ctypes.byref(b), ctypes.byref(f)))
        return b.value, f.value

    def s4_kernels(self):
        n = ctypes.c_int(0)
        check(lib().dwb_plan_s4_blocks(self._plan, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            H, l = ctypes.c_int(0), ctypes.c_int(0)
            check(lib().dwb_plan_s4_kernel(self._plan, i, None, 0, ctypes.byref(H), ctypes.byref(l)))
            k = torch.empty(2, H.value, l.value, dtype=torch.float32, device=self.device)
            check(lib().dwb_plan_s4_kernel(self._plan, i, ptr(k), k.numel(), ctypes.byref(H), ctypes.byref(l)))
            out.append(k)
        return out

    def close(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            lib().dwb_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
