"""The reference's plugin surface (`models/__init__.py:4-23`) backed by the libdwb engine.

`construct_model(cfg)` returns an `nn.Module` whose parameters carry the reference's exact
state_dict keys and shapes, so `load_state_dict(checkpoint['model_state_dict'])`
works unchanged.  The modules are parameter containers: `forward((audio, steps), mel_spec)` hands
the tensors to the CUDA engine (`engine.Engine`).  There is deliberately no PyTorch or CPU
execution path and no autograd: training goes through `training.Trainer` (native backward + Adam,
WaveNet only), so a forward under autograd, without CUDA tensors, or with the
native library missing, raises.
"""
import math

import torch
import torch.nn as nn

from . import init as _init


class _WNConv(nn.Module):
    """Weight-normed Conv1d holder registered as `.conv` (reference `Conv`, wavenet.py:16-26)."""

    def __init__(self, cin, cout, k=1):
        super().__init__()
        self.conv = _wn_conv1d(cin, cout, k)


class _WNParams(nn.Module):
    """weight_g / weight_v / bias with PyTorch's default conv init and g = ||v|| (SURVEY Appendix E:
    the reference's kaiming call after weight_norm only touches a derived tensor)."""

    def __init__(self, shape_v, fan_in, norm_dims):
        super().__init__()
        bound = 1.0 / math.sqrt(fan_in)
        v = torch.empty(shape_v).uniform_(-bound, bound)
        self.bias = nn.Parameter(torch.empty(shape_v[0] if norm_dims != "all" else shape_v[1]).uniform_(-bound, bound))
        if norm_dims == "all":
            g = v.norm().reshape(1, 1, 1, 1)
        else:
            g = v.reshape(shape_v[0], -1).norm(dim=1).reshape((-1,) + (1,) * (len(shape_v) - 1))
        self.weight_g = nn.Parameter(g)
        self.weight_v = nn.Parameter(v)


def _wn_conv1d(cin, cout, k):
    return _WNParams((cout, cin, k), cin * k, "rows")


def _wn_convT2d(s):
    # ConvTranspose2d(1, 1, (3, 2s)): weight (1,1,3,2s); dim 0 has size 1 -> a single scalar norm
    return _WNParams((1, 1, 3, 2 * s), 3 * 2 * s, "all")


class _Plain1x1(nn.Module):
    def __init__(self, cin, cout, zero=False):
        super().__init__()
        bound = 1.0 / math.sqrt(cin)
        self.weight = nn.Parameter(torch.zeros(cout, cin, 1) if zero else torch.empty(cout, cin, 1).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.zeros(cout) if zero else torch.empty(cout).uniform_(-bound, bound))


class _ZeroConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _Plain1x1(cin, cout, zero=True)


class _Linear(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        bound = 1.0 / math.sqrt(cin)
        self.weight = nn.Parameter(torch.empty(cout, cin).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


def _mel_parts(mod, channels, mel_upsample):
    mod.upsample_conv2d = nn.ModuleList([_wn_convT2d(s) for s in mel_upsample])
    mod.mel_conv = _WNConv(80, channels, 1)


class _EngineBacked(nn.Module):
    """Shared dispatch: build the CUDA plan lazily from the current parameters, run it."""

    _cfg: dict

    def _fingerprint(self):
        # (address, in-place version) of every parameter and buffer: changes under optimizer steps, submodule
        # load_state_dict, .to()/.cuda(), in-place edits.  Writes through `p.data` do not bump the version counter:
        # call invalidate() after those.
        ts = self.__dict__.get("_fp_tensors")
        if ts is None:                  # the Parameter / buffer objects themselves are stable: cache the traversal
            ts = self.__dict__["_fp_tensors"] = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _engine_get(self):
        from .engine import Engine
        eng = self.__dict__.get("_engine")
        if eng is not None and self.__dict__.get("_engine_fp") != self._fingerprint():
            self.invalidate()           # parameters changed under the compiled plan: rebuild rather than serve stale weights
            eng = None
        if eng is None:
            eng = Engine(self._cfg, self)
            self.__dict__["_engine"] = eng
            self.__dict__["_engine_fp"] = self._fingerprint()     # after the build: it may rewrite fresh S4 kernels in place
        return eng

    def invalidate(self):
        """Drop the compiled plan (needed only after writes through `p.data`; everything else is detected)."""
        eng = self.__dict__.pop("_engine", None)
        self.__dict__.pop("_engine_fp", None)
        self.__dict__.pop("_fp_tensors", None)
        if eng is not None:
            eng.close()

    def __getstate__(self):             # the plan is a ctypes handle: never pickled or deep-copied with the module
        st = self.__dict__.copy()
        st.pop("_engine", None)
        st.pop("_engine_fp", None)
        st.pop("_fp_tensors", None)
        return st

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def forward(self, input_data, mel_spec=None):
        audio, diffusion_steps = input_data
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError(
                "diffwave_sashimi_b200 modules run the inference engine only: call under torch.no_grad(). "
                "There is no autograd path; training runs through diffwave_sashimi_b200.training.Trainer "
                "(native backward, model=wavenet).")
        if not audio.is_cuda:
            raise RuntimeError("diffwave_sashimi_b200 has no CPU path: move the model and inputs to a B200 (.cuda())")
        return self._engine_get().forward(audio, diffusion_steps, mel_spec)


# ----------------------------------------------------------------------------------------
class _ResidualBlock(nn.Module):
    def __init__(self, C, S, E_out, unconditional, mel_upsample):
        super().__init__()
        self.fc_t = _Linear(E_out, C)
        self.dilated_conv_layer = _WNConv(C, 2 * C, 3)
        if not unconditional:
            _mel_parts(self, 2 * C, mel_upsample)
        self.res_conv = _wn_conv1d(C, C, 1)
        self.skip_conv = _wn_conv1d(C, S, 1)


class _ResidualGroup(nn.Module):
    def __init__(self, C, S, N, E_in, E_mid, E_out, unconditional, mel_upsample):
        super().__init__()
        self.fc_t1 = _Linear(E_in, E_mid)
        self.fc_t2 = _Linear(E_mid, E_out)
        self.residual_blocks = nn.ModuleList(
            [_ResidualBlock(C, S, E_out, unconditional, mel_upsample) for _ in range(N)])


class WaveNet(_EngineBacked):
    """models/wavenet.py:168-220 (same ctor kwargs, same parameter names)."""

    def __init__(self, in_channels=1, res_channels=256, skip_channels=128, out_channels=1,
                 num_res_layers=30, dilation_cycle=10,
                 diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                 diffusion_step_embed_dim_out=512, unconditional=False, mel_upsample=[16, 16], **kwargs):
        super().__init__()
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("the engine implements mono audio (in_channels = out_channels = 1)")
        self.res_channels, self.skip_channels = res_channels, skip_channels
        self.num_res_layers, self.unconditional = num_res_layers, unconditional
        self._cfg = dict(_name_="wavenet", unconditional=unconditional, res_channels=res_channels,
                         skip_channels=skip_channels, num_res_layers=num_res_layers, dilation_cycle=dilation_cycle,
                         diffusion_step_embed_dim_in=diffusion_step_embed_dim_in,
                         diffusion_step_embed_dim_mid=diffusion_step_embed_dim_mid,
                         diffusion_step_embed_dim_out=diffusion_step_embed_dim_out, mel_upsample=list(mel_upsample))
        self.init_conv = nn.Sequential(_WNConv(in_channels, res_channels, 1), nn.ReLU())
        self.residual_layer = _ResidualGroup(res_channels, skip_channels, num_res_layers,
                                             diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid,
                                             diffusion_step_embed_dim_out, unconditional, mel_upsample)
        self.final_conv = nn.Sequential(_WNConv(skip_channels, skip_channels, 1), nn.ReLU(),
                                        _ZeroConv(skip_channels, out_channels))

    def __repr__(self):
        return f"wavenet_h{self.res_channels}_d{self.num_res_layers}_{'uncond' if self.unconditional else 'cond'}"

    @classmethod
    def name(cls, cfg):
        # the reference's version raises NameError (wavenet.py:215-220 uses an undefined `model_cfg`);
        # this is the value its exp/ directories were created with
        return "wnet_h{}_d{}".format(cfg["res_channels"], cfg["num_res_layers"])


# ----------------------------------------------------------------------------------------
class _TLN(nn.Module):
    def __init__(self):
        super().__init__()
        self.m = nn.Parameter(torch.zeros(1))
        self.s = nn.Parameter(torch.ones(1))


class _SSKernelParams(nn.Module):
    def __init__(self, H, d_state):
        super().__init__()
        p = _init.s4_layer_params(H, d_state)
        for k in ("C", "log_dt", "B", "P", "inv_w_real", "w_imag"):
            self.register_parameter(k, nn.Parameter(p["kernel.kernel." + k]))
        self.register_buffer("L", p["kernel.kernel.L"])
        self._D = p["D"]


class _SSKernel(nn.Module):
    def __init__(self, H, d_state):
        super().__init__()
        self.kernel = _SSKernelParams(H, d_state)


class _S4(nn.Module):
    def __init__(self, H, l_max, d_state=64):
        super().__init__()
        self.H, self.l_max = H, l_max
        self.kernel = _SSKernel(H, d_state)
        self.D = nn.Parameter(self.kernel.kernel.__dict__.pop("_D"))
        self.output_linear = nn.Sequential(_Plain1x1(H, 2 * H), nn.GLU(dim=-2))


class _FF(nn.Module):
    def __init__(self, H, expand):
        super().__init__()
        self.ff = nn.Sequential(_WNConv(H, expand * H, 1), nn.GELU(), _WNConv(expand * H, H, 1))


class _DiffWaveBlock(nn.Module):
    def __init__(self, H, l, ff, unconditional, mel_upsample, E_out=512):
        super().__init__()
        self.fc_t = _Linear(E_out, H)      # the reference hard-wires 512 here (sashimi.py:116)
        self.layer = _S4(H, l)
        self.ff = _FF(H, ff)
        self.norm1, self.norm2 = _TLN(), _TLN()
        if not unconditional:
            _mel_parts(self, H, mel_upsample)


class _Pool(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear = _WNConv(cin, cout, 1)


class Sashimi(_EngineBacked):
    """models/sashimi.py:188-327 (same ctor kwargs, same parameter names)."""

    def __init__(self, in_channels=1, out_channels=1, d_model=64, n_layers=8, pool=[4, 4], expand=2, ff=2,
                 unet=True, diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                 diffusion_step_embed_dim_out=512, unconditional=False, mel_upsample=[16, 16], L=16000, **kwargs):
        super().__init__()
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("the engine implements mono audio (in_channels = out_channels = 1)")
        if diffusion_step_embed_dim_out != 512:
            raise ValueError("the reference's DiffWaveBlock fixes diffusion_step_embed_dim_out at 512 (sashimi.py:116)")
        self.L, self.unet, self.d_model, self.n_layers = L, unet, d_model, n_layers
        self.expand, self.ff, self.pool, self.unconditional = expand, ff, list(pool), unconditional
        self._cfg = dict(_name_="sashimi", unconditional=unconditional, d_model=d_model, n_layers=n_layers,
                         pool=list(pool), expand=expand, ff=ff, unet=unet, L=L,
                         diffusion_step_embed_dim_in=diffusion_step_embed_dim_in,
                         diffusion_step_embed_dim_mid=diffusion_step_embed_dim_mid,
                         diffusion_step_embed_dim_out=diffusion_step_embed_dim_out, mel_upsample=list(mel_upsample))
        self.init_conv = nn.Sequential(_WNConv(in_channels, d_model, 1), nn.ReLU())
        self.fc_t1 = _Linear(diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid)
        self.fc_t2 = _Linear(diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out)
        blk = lambda H, l: _DiffWaveBlock(H, l, ff, unconditional, mel_upsample)
        H, l = d_model, L
        d_layers = []
        for p in pool:
            if unet:
                d_layers += [blk(H, l) for _ in range(n_layers)]
            d_layers.append(_Pool(H * p, H * expand))
            l //= p
            H *= expand
        self.d_layers = nn.ModuleList(d_layers)
        self.c_layers = nn.ModuleList([blk(H, l) for _ in range(n_layers)])
        u_layers = []
        for p in pool[::-1]:
            H //= expand
            l *= p
            u_layers.append(_Pool(H * expand, H * p))
            u_layers += [blk(H, l) for _ in range(n_layers)]
        self.u_layers = nn.ModuleList(u_layers)
        self.norm = _TLN()
        self.final_conv = nn.Sequential(_WNConv(d_model, d_model, 1), nn.ReLU(), _ZeroConv(d_model, out_channels))

    def __repr__(self):
        # the reference's __repr__ raises (sashimi.py:315-316); this is its evident intent
        return (f"sashimi_h{self.d_model}_d{self.n_layers}_pool{''.join(map(str, self.pool))}_expand{self.expand}"
                f"_ff{self.ff}_{'uncond' if self.unconditional else 'cond'}")

    @classmethod
    def name(cls, cfg):
        return "{}_d{}_n{}_pool_{}_expand{}_ff{}".format(
            "unet" if cfg["unet"] else "snet", cfg["d_model"], cfg["n_layers"], len(cfg["pool"]),
            cfg["expand"], cfg["ff"])


_REGISTRY = {"wavenet": WaveNet, "sashimi": Sashimi}


def construct_model(model_cfg):
    """models/__init__.py:4-12: pop `_name_`, build, restore."""
    name = model_cfg.pop("_name_")
    try:
        model = _REGISTRY[name](**model_cfg)
    finally:
        model_cfg["_name_"] = name
    return model


def model_identifier(model_cfg):
    """models/__init__.py:18-23."""
    name = model_cfg["_name_"] if isinstance(model_cfg, dict) else model_cfg._name_
    return _REGISTRY[name].name(model_cfg)
