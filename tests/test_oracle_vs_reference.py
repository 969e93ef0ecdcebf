"""Oracle vs the LIVE reference (CPU): only where /root/reference is mounted (the build container); skipped on
the GPU box, where the committed goldens of tests/test_oracle_vs_golden.py carry the same pinning.  Seeds and
shapes here differ from the goldens', so this is an independent sample of the same claim."""
import os

import pytest
import torch

from conftest import rel_l2, rel_max
from oracle import diffwave_oracle as O
from oracle import refshim

pytestmark = pytest.mark.skipif(not os.path.isdir(refshim.REF_ROOT), reason="reference not mounted")

EMB = dict(diffusion_step_embed_dim_in=16, diffusion_step_embed_dim_mid=32, diffusion_step_embed_dim_out=512)
CASES = [
    ("unet_d64", dict(d_model=8, n_layers=1, L=192, pool=[4, 2], **EMB), 192, None),
    ("unet_d64", dict(d_model=8, n_layers=1, L=192, pool=[4, 2], **EMB), 96, None),        # L < l_max: kernel truncation
    ("unet_d64", dict(d_model=8, n_layers=1, L=192, pool=[4, 2], **EMB), 480, None),       # L > l_max
    ("unet_d32_cond", dict(d_model=8, n_layers=1, L=512, **EMB), 512, (1, 80, 3)),
    ("wnet_h128_d30", dict(res_channels=8, skip_channels=8, num_res_layers=4, dilation_cycle=3,
                           diffusion_step_embed_dim_in=16, diffusion_step_embed_dim_mid=32, diffusion_step_embed_dim_out=32), 200, None),
]


@pytest.mark.parametrize("base,over,Lr,melshape", CASES)
def test_forward_matches_live_reference(base, over, Lr, melshape):
    ns = refshim.load(parity=True)
    cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
    cfg.update(over)
    torch.manual_seed(21)
    net = ns.models.construct_model(cfg).eval()
    with torch.no_grad():
        w = net.final_conv[2].conv.weight
        w.normal_(0, (1.0 / w.shape[1]) ** 0.5)
    g = torch.Generator().manual_seed(22)
    x = torch.randn(2, 1, Lr, generator=g)
    t = torch.tensor([[4.0], [31.0]])
    mel = torch.randn(*melshape, generator=g) if melshape else None
    with torch.no_grad():
        if cfg["_name_"] == "sashimi":        # settle the one-off C rewrite at the configured length first
            net((torch.zeros(1, 1, cfg["L"]), torch.zeros(1, 1)), mel_spec=None if mel is None else torch.zeros(1, 80, cfg["L"] // 256 + 1))
        ref = net((x, t), mel_spec=mel)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    got = O.forward(dict(cfg), sd, x, t, mel=mel)
    assert rel_l2(got, ref) < 2e-5 and rel_max(got, ref) < 2e-5
