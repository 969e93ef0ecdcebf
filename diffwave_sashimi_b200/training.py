"""Host mirror of the training step (train.py:84-143,198-222) over libdwb's trainer entries.

    net = dwb.construct_model(model_cfg).cuda()
    trainer = dwb.training.Trainer(net, batch_size, audio_length, lr=2e-4)     # ~ torch.optim.Adam(net.parameters(), lr)
    loss = trainer.loss_backward(audio, diffusion_hyperparams)                 # ~ zero_grad(); training_loss(); backward()
    trainer.allreduce_gradients()                                              # ~ apply_gradient_allreduce's hook
    trainer.step()                                                             # ~ optimizer.step()

All arithmetic (forward with saved activations, every gradient, Adam) runs in libdwb kernels; there is no autograd
and no PyTorch fallback: without the library or a GPU the constructor raises.  Parameters, gradients and the Adam
moments live in four flat buffers in net.parameters() order; every nn.Parameter of `net` becomes a view of the flat
parameter buffer and its .grad a view of the flat gradient buffer, so `net.state_dict()`, checkpoints in the
reference's format and `p.grad` inspection keep working, while the gradient exchange is a single all-reduce.

Built for model._name_ = wavenet, unconditional (configs[0], configs[4] of BASELINE.json).  SaShiMi and
mel-conditioned training raise DwbError(DWB_ERR_UNSUPPORTED): their backward kernels are not written.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .engine import config_struct


def trainer_layout(cfg: dict):
    """[(state_dict key, float offset, numel)] in net.parameters() order, and the total float count.
    Pure host call (works without a GPU)."""
    c = config_struct(cfg)
    n, total = ctypes.c_int(0), ctypes.c_int64(0)
    check(lib().dwb_trainer_layout(ctypes.byref(c), -1, None, 0, None, None, ctypes.byref(n), ctypes.byref(total)))
    out = []
    buf = ctypes.create_string_buffer(256)
    for i in range(n.value):
        off, numel = ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib().dwb_trainer_layout(ctypes.byref(c), i, buf, 256, ctypes.byref(off), ctypes.byref(numel), None, None))
        out.append((buf.value.decode(), off.value, numel.value))
    return out, total.value


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, betas=(0.9, 0.999), eps=1e-8, step=1, grad_scale=1.0):
    """torch.optim.Adam.step() on flat CUDA f32 buffers, one launch (dwb_adam_step)."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()):
            raise ValueError("adam_step needs four contiguous CUDA float32 buffers of equal size")
    check(lib().dwb_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(), lr, betas[0], betas[1],
                              eps, step, grad_scale, stream_ptr(params.device)))


class Trainer:
    """Optimizer + backward of one model for batches of exactly (batch_size, 1, audio_length)."""

    def __init__(self, net, batch_size, audio_length, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, gemm=None):
        """gemm: None = the library default (DWB_TRAIN_GEMM), "mma" = split-bf16 tensor cores, "simt" = exact fp32 tiles."""
        cfg = dict(net._cfg)
        named = list(net.named_parameters())
        if not named or not named[0][1].is_cuda:
            raise RuntimeError("diffwave_sashimi_b200 has no CPU path: move the model to a B200 (.cuda()) before building a Trainer")
        self.device = named[0][1].device
        self.net, self.cfg, self.B, self.L = net, cfg, int(batch_size), int(audio_length)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.layout, total = trainer_layout(cfg)
        if [(k, p.numel()) for k, p in named] != [(k, n) for k, _, n in self.layout]:
            raise RuntimeError("parameter order of the module differs from dwb_trainer_layout")
        self.params = torch.empty(total, device=self.device, dtype=torch.float32)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        with torch.no_grad():
            for (_, p), (_, off, n) in zip(named, self.layout):
                view = self.params[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                                  # the module now reads and writes the flat buffer
                p.grad = self.grads[off:off + n].view(p.shape)
        self._first, self._last = named[0][1], named[-1][1]
        self.n_steps = 0
        self.world_size = 1
        self._scale = 1.0
        self._loss = torch.zeros((), device=self.device, dtype=torch.float32)
        self._h = ctypes.c_void_p()
        c = config_struct(cfg)
        with torch.cuda.device(self.device):
            check(lib().dwb_trainer_create(ctypes.byref(c), self.device.index or 0, self.B, self.L, ctypes.byref(self._h)))
        if gemm is not None:
            if gemm not in ("mma", "simt"):
                raise ValueError("gemm must be None, 'mma' or 'simt'")
            check(lib().dwb_trainer_set_gemm(self._h, int(gemm == "mma")))

    def _check_bound(self):
        """The module's parameters must still be the views of the flat buffer made in __init__: `.cuda()` / `.to()` /
        `p.data = ...` after construction would re-home them, and training would silently update memory the module no
        longer reads."""
        for (_, off, _), p in ((self.layout[0], self._first), (self.layout[-1], self._last)):
            if p.data_ptr() != self.params.data_ptr() + 4 * off:
                raise RuntimeError("the model's parameters were moved after Trainer(net, ...) was built (.cuda()/.to()/p.data=...): "
                                   "build the Trainer after the last move")

    # ---- loss + backward (train.py:198-222 + loss.backward()) ---------------------------------------------------
    def loss_backward(self, audio, diffusion_hyperparams, mel_spec=None, diffusion_steps=None, z=None, return_eps=False):
        """Gradients are overwritten (an implicit optimizer.zero_grad()).  Returns the loss as a 0-dim CUDA tensor
        (a fresh tensor per call), or (loss, eps) with return_eps.  `diffusion_steps` (B,) / `z` default to the
        reference's CPU-generator draws, in its order (train.py:217-218)."""
        if mel_spec is not None:
            raise _lib.DwbError(5, "mel-conditioned training is not implemented (unconditional WaveNet only)")
        self._check_bound()
        B, C, L = audio.shape
        if (B, C, L) != (self.B, 1, self.L) or not audio.is_cuda:
            raise ValueError(f"this Trainer was built for CUDA batches of shape ({self.B}, 1, {self.L}), got {tuple(audio.shape)}")
        T, alpha_bar = diffusion_hyperparams["T"], diffusion_hyperparams["Alpha_bar"]
        if diffusion_steps is None:
            diffusion_steps = torch.randint(T, size=(B, 1, 1))
        if z is None:
            z = torch.normal(0, 1, size=audio.shape)
        steps = diffusion_steps.reshape(B).to(self.device)
        z = z.to(self.device, torch.float32).contiguous()
        ab = alpha_bar.to(self.device, torch.float32)[steps.long()]
        coef = torch.stack([torch.sqrt(ab), torch.sqrt(1 - ab)], dim=1).contiguous()      # same fp32 torch ops as train.py:219
        audio = audio.to(torch.float32).contiguous()
        eps = torch.empty_like(audio) if return_eps else None
        with torch.cuda.device(self.device):
            check(lib().dwb_trainer_loss_backward(self._h, ptr(self.params), ptr(self.grads), ptr(audio), ptr(z),
                                                  ptr(steps.to(torch.float32).contiguous()), ptr(coef), ptr(eps), ptr(self._loss),
                                                  stream_ptr(self.device)))
        self._scale = 1.0
        loss = self._loss.clone()
        return (loss, eps) if return_eps else loss

    # ---- data parallel (distributed_util.py:97-149) ---------------------------------------------------------------
    def broadcast_parameters(self, src=0):
        from .distributed import broadcast_flat
        broadcast_flat(self.params, src)

    def allreduce_gradients(self):
        """Sum the flat gradient buffer over ranks; the 1/world_size of the reference's `coalesced /= world_size`
        is folded into the Adam kernel (grads read back through p.grad between this call and step() are sums)."""
        from .distributed import allreduce_flat
        self.world_size = allreduce_flat(self.grads, average=False)
        self._scale = 1.0 / self.world_size

    # ---- optimizer.step() ---------------------------------------------------------------------------------------------
    def step(self):
        self._check_bound()
        self.n_steps += 1
        with torch.cuda.device(self.device):
            adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.lr, self.betas, self.eps, self.n_steps, self._scale)
        if hasattr(self.net, "invalidate"):
            self.net.invalidate()               # the inference plan holds folded copies of the old weights

    def zero_grad(self):
        """Kept for loop compatibility: loss_backward overwrites every gradient."""

    # ---- torch.optim.Adam-compatible state (train.py:158-160,101-105) ---------------------------------------------
    def state_dict(self):
        state = {}
        for i, (_, off, n) in enumerate(self.layout):
            shape = self._shape(i)
            state[i] = {"step": torch.tensor(float(self.n_steps)), "exp_avg": self.exp_avg[off:off + n].view(shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view(shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.layout)))}
        return {"state": state if self.n_steps else {}, "param_groups": [group]}

    def load_state_dict(self, sd):
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.n_steps = 0
        for i, st in sd.get("state", {}).items():            # parameters that never had a gradient have no entry
            _, off, n = self.layout[int(i)]
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.n_steps = max(self.n_steps, int(float(st["step"])))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"])

    def _shape(self, i):
        return dict(self.net.named_parameters())[self.layout[i][0]].shape

    # ---- introspection --------------------------------------------------------------------------------------------------
    def info(self):
        ws, n = ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib().dwb_trainer_info(self._h, ctypes.byref(ws), ctypes.byref(n)))
        return {"workspace_bytes": ws.value, "launches": n.value, "parameters": self.params.numel()}

    def close(self):
        if self._h:
            lib().dwb_trainer_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
