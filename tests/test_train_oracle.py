"""CPU: the training-step oracle (oracle/train_oracle.py) against fixtures made from the unmodified reference
(tests/golden/make_golden.py --train: reference modules + torch autograd + torch.optim.Adam, train.py:84-143,198-222),
and the C ABI's parameter layout against the module's net.parameters() order."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as TO

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    cfg = ast.literal_eval(str(z["cfg"]))
    return z, cfg


def sub(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def alpha_bar(z):
    from diffwave_sashimi_b200.sampler import calc_diffusion_hyperparams
    return calc_diffusion_hyperparams(int(z["T"]), float(z["beta_0"]), float(z["beta_T"]))["Alpha_bar"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["train_wnet_a", "train_wnet_b"])
@pytest.mark.parametrize("form", ["autograd", "manual"])
def test_loss_and_gradients_vs_reference(name, form):
    z, cfg = load(name)
    sd0, ref = sub(z, "sd0/"), sub(z, "grad0/")
    fn = TO.loss_and_grads_autograd if form == "autograd" else TO.loss_and_grads_manual
    loss, eps, grads = fn(cfg, sd0, torch.from_numpy(z["audio0"]), torch.from_numpy(z["steps0"]), torch.from_numpy(z["z0"]), alpha_bar(z))
    assert abs(float(loss) - z["losses"][0]) <= 2e-6 * z["losses"][0]
    assert rel(eps, torch.from_numpy(z["eps0"])) < 2e-6
    assert set(ref) <= set(grads)
    gscale = max(float(g.norm()) for g in ref.values())
    for k, g in ref.items():
        # the reference leaves .grad = None where the graph does not reach (last layer's res_conv): stored as zeros
        err = float((grads[k].double() - g.double()).norm())
        assert err <= 2e-5 * float(g.norm()) + 1e-7 * gscale, (k, err, float(g.norm()))


def test_sashimi_loss_and_gradients_vs_reference():
    """The oracle for the half of the training row without kernels: loss, eps and all 236 parameter gradients of a tiny
    SaShiMi UNet - through the per-step S4 kernel generation, down to C, B, P, inv_w_real, w_imag, log_dt - against the
    reference's own autograd (models/s4.py:674-807, models/sashimi.py:143-184)."""
    z, cfg = load("train_unet_tiny")
    w = np.load(os.path.join(GOLD, "tiny_unet.npz"))
    sd = {k[3:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("sd/")}
    ref = sub(z, "grad0/")
    loss, eps, grads = TO.loss_and_grads_autograd(cfg, sd, torch.from_numpy(z["audio0"]), torch.from_numpy(z["steps0"]),
                                                  torch.from_numpy(z["z0"]), alpha_bar(z))
    assert abs(float(loss) - z["losses"][0]) <= 1e-5 * z["losses"][0]
    assert rel(eps, torch.from_numpy(z["eps0"])) < 1e-5
    gscale = max(float(g.norm()) for g in ref.values())
    assert len(ref) == 236 and any(k.endswith("kernel.kernel.log_dt") for k in ref)
    for k, g in ref.items():
        err = float((grads[k].double() - g.double()).norm())
        assert err <= 1e-3 * float(g.norm()) + 1e-6 * gscale, (k, err, float(g.norm()))


def test_manual_backward_equals_autograd_fp64():
    z, cfg = load("train_wnet_b")
    args = (cfg, sub(z, "sd0/"), torch.from_numpy(z["audio1"]), torch.from_numpy(z["steps1"]), torch.from_numpy(z["z1"]), alpha_bar(z))
    la, _, ga = TO.loss_and_grads_autograd(*args)
    lm, _, gm = TO.loss_and_grads_manual(*args)
    assert abs(float(la) - float(lm)) < 1e-12
    for k in gm:
        assert float((ga[k] - gm[k]).norm()) <= 1e-10 * float(ga[k].norm()) + 1e-14, k


@pytest.mark.parametrize("name", ["train_wnet_a", "train_wnet_b"])
def test_three_adam_steps_vs_reference(name):
    z, cfg = load(name)
    sd = {k: v.double() for k, v in sub(z, "sd0/").items()}
    ab = alpha_bar(z)
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    has = {k[len("hasgrad/"):]: bool(z[k]) for k in z.files if k.startswith("hasgrad/")}
    for it in range(3):
        loss, _, grads = TO.loss_and_grads_manual(cfg, sd, torch.from_numpy(z[f"audio{it}"]), torch.from_numpy(z[f"steps{it}"]),
                                                  torch.from_numpy(z[f"z{it}"]), ab)
        assert abs(float(loss) - z["losses"][it]) <= 5e-6 * z["losses"][it]
        for k in sd:
            if has[k]:
                TO.adam_step(sd[k], grads[k], m[k], v2[k], float(z["lr"]), it + 1)
    sd3, sd0, g0 = sub(z, "sd3/"), sub(z, "sd0/"), sub(z, "grad0/")
    gscale = max(float(g.norm()) for g in g0.values())
    for k in sd:
        if has[k] and float(g0[k].norm()) < 1e-6 * gscale:
            continue        # d/dv of g v/|v| on a one-element row is exactly 0: the reference's Adam amplifies fp32 rounding noise there
        moved = (sd3[k] - sd0[k]).double()
        assert float((sd[k] - sd0[k].double() - moved).norm()) <= 2e-3 * float(moved.norm()) + 1e-9, k


def test_parameter_layout_matches_module():
    import diffwave_sashimi_b200 as dwb
    from diffwave_sashimi_b200.training import trainer_layout
    for over in (dict(res_channels=16, skip_channels=8, num_res_layers=3, dilation_cycle=2),
                 dict(res_channels=128, skip_channels=256, num_res_layers=30, dilation_cycle=10)):
        cfg = dict(_name_="wavenet", unconditional=True, **over)
        net = dwb.construct_model(dict(cfg))
        lay, total = trainer_layout(cfg)
        named = list(net.named_parameters())
        assert [k for k, _, _ in lay] == [k for k, _ in named]
        assert [n for _, _, n in lay] == [p.numel() for _, p in named]
        assert total == sum(p.numel() for _, p in named) and lay[0][1] == 0
        assert all(lay[i][1] + lay[i][2] == lay[i + 1][1] for i in range(len(lay) - 1))


def test_trainer_rejects_what_is_not_built():
    from diffwave_sashimi_b200._lib import DwbError
    from diffwave_sashimi_b200.training import trainer_layout
    with pytest.raises(DwbError) as e:
        trainer_layout(dict(_name_="sashimi", unconditional=True, d_model=64, n_layers=6, pool=[4, 4], expand=2, ff=2, L=16000))
    assert e.value.code == 5
    with pytest.raises(DwbError):
        trainer_layout(dict(_name_="wavenet", unconditional=False, res_channels=16, skip_channels=8, num_res_layers=3, dilation_cycle=2))


def test_trainer_needs_a_gpu_model():
    """No CPU path: a Trainer over CPU parameters raises before anything native is created."""
    import diffwave_sashimi_b200 as dwb
    from diffwave_sashimi_b200.training import Trainer
    net = dwb.construct_model(dict(_name_="wavenet", unconditional=True, res_channels=16, skip_channels=8, num_res_layers=2, dilation_cycle=2))
    with pytest.raises(RuntimeError, match="no CPU path"):
        Trainer(net, 2, 64)
