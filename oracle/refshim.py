"""TEST INFRASTRUCTURE — import shim for the *unmodified* reference at /root/reference.

Only used (a) by tests/golden/make_golden.py to generate committed fixtures and
(b) by `-m "not gpu"` tests that compare oracle/ against the live reference when the
mount exists (this container).  Nothing on the product path imports this file, and
nothing here runs on the GPU box (where /root/reference does not exist).

Recipe = SURVEY.md Appendix C: stub the four uninstalled modules, make `.cuda()` the
identity, and (parity mode) replace the reference's CPU Cauchy fallback
(models/s4.py:109-116, which drops the conjugate pair) with the conjugate-symmetric sum
its CUDA kernel computes (extensions/cauchy/cauchy_cuda.cu:331).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DWB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "s4.py"))


class Cfg(dict):
    """dict with attribute access (construct_model uses .pop/[]; model_identifier uses ._name_)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


_loaded = {}


def load(parity: bool = True):
    """Return a namespace with the reference's `models`, `generate.sampling`, `utils`.

    parity=True  -> models.s4.cauchy_naive patched to the conjugate-symmetric form.
    parity=False -> as shipped (timing only; computes a different kernel, SURVEY finding 3).
    """
    import torch

    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    if "ns" not in _loaded:
        pl = types.ModuleType("pytorch_lightning")
        plu = types.ModuleType("pytorch_lightning.utilities")
        plu.rank_zero_only = lambda f: f
        pl.utilities = plu
        sys.modules.setdefault("pytorch_lightning", pl)
        sys.modules.setdefault("pytorch_lightning.utilities", plu)

        oe = types.ModuleType("opt_einsum")

        def contract(expr, *ops, **kw):
            return torch.einsum(expr.replace(" ", ""), *ops)

        def contract_expression(expr, *shapes, **kw):
            return lambda *ops: torch.einsum(expr.replace(" ", ""), *ops)

        oe.contract, oe.contract_expression = contract, contract_expression
        sys.modules.setdefault("opt_einsum", oe)

        hy = types.ModuleType("hydra")
        hy.main = lambda *a, **k: (lambda f: f)
        sys.modules.setdefault("hydra", hy)
        om = types.ModuleType("omegaconf")
        om.DictConfig, om.OmegaConf = dict, type("OmegaConf", (), {})
        sys.modules.setdefault("omegaconf", om)

        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

        sys.path.insert(0, REF_ROOT)
        import logging
        logging.disable(logging.ERROR)
        import models as ref_models  # noqa
        import models.s4 as ref_s4  # noqa
        import generate as ref_generate  # noqa
        import utils as ref_utils  # noqa
        logging.disable(logging.NOTSET)
        sys.path.remove(REF_ROOT)

        ns = types.SimpleNamespace(models=ref_models, s4=ref_s4, generate=ref_generate,
                                   utils=ref_utils, shipped_cauchy=ref_s4.cauchy_naive)
        _loaded["ns"] = ns
    ns = _loaded["ns"]

    def cauchy_sym(v, z, w):
        # sum_n v/(z-w) + conj(v)/(z-conj(w))   (cauchy_cuda.cu:331 semantics)
        zz = z.unsqueeze(-2)
        vv, ww = v.unsqueeze(-1), w.unsqueeze(-1)
        return (vv / (zz - ww) + vv.conj() / (zz - ww.conj())).sum(dim=-2)

    ns.s4.cauchy_naive = cauchy_sym if parity else ns.shipped_cauchy
    return ns


MODEL_CFGS = {
    # configs/model/wavenet_small.yaml, wavenet.yaml, sashimi_small.yaml, sashimi.yaml
    "wnet_h128_d30": dict(_name_="wavenet", unconditional=True, in_channels=1, out_channels=1,
                          diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                          diffusion_step_embed_dim_out=512, res_channels=128, skip_channels=256,
                          num_res_layers=30, dilation_cycle=10),
    "wnet_h256_d36": dict(_name_="wavenet", unconditional=True, in_channels=1, out_channels=1,
                          diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                          diffusion_step_embed_dim_out=512, res_channels=256, skip_channels=256,
                          num_res_layers=36, dilation_cycle=12),
    "unet_d64": dict(_name_="sashimi", unconditional=True, in_channels=1, out_channels=1,
                     diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                     diffusion_step_embed_dim_out=512, unet=True, d_model=64, n_layers=6,
                     pool=[4, 4], expand=2, ff=2, L=16000),
    "unet_d128": dict(_name_="sashimi", unconditional=True, in_channels=1, out_channels=1,
                      diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                      diffusion_step_embed_dim_out=512, unet=True, d_model=128, n_layers=6,
                      pool=[4, 4], expand=2, ff=2, L=16000),
    "unet_d32_cond": dict(_name_="sashimi", unconditional=False, in_channels=1, out_channels=1,
                          diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                          diffusion_step_embed_dim_out=512, unet=True, d_model=32, n_layers=6,
                          pool=[4, 4], expand=2, ff=2, L=16000, mel_upsample=[16, 16]),
}
