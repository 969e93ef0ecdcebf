"""`-m gpu`: the native training step (loss + every gradient + Adam; csrc/train_wavenet.cu through the C ABI) against
fixtures from the unmodified reference modules + torch autograd + torch.optim.Adam (tests/golden/make_golden.py --train;
train.py:84-143,198-222) and against the fp64 oracle (oracle/train_oracle.py)."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu

# exact-fp32 SIMT tiles against fp32 autograd: measured ~1e-6 (printed); split-bf16 tensor-core GEMMs (three MMAs per
# product, ~2^-17 each): measured values printed, bars well inside the north star's 1e-3
GRAD_TOL = {"simt": 1e-4, "mma": 5e-4}
LOSS_TOL = {"simt": 1e-5, "mma": 1e-4}
GEMMS = ["simt", "mma"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return z, ast.literal_eval(str(z["cfg"]))


def sub(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


@pytest.fixture(scope="module")
def dwb():
    import diffwave_sashimi_b200 as d
    assert torch.cuda.is_available()
    return d


def _trainer(dwb, cfg, sd, B, L, **kw):
    from diffwave_sashimi_b200.training import Trainer
    net = dwb.construct_model(dict(cfg))
    net.load_state_dict(sd)
    net = net.cuda().train()
    return net, Trainer(net, B, L, **kw)


def _dh(dwb, z):
    return dwb.calc_diffusion_hyperparams(int(z["T"]), float(z["beta_0"]), float(z["beta_T"]))


def _check_grads(net, ref, label, gemm="simt"):
    gscale = max(float(g.norm()) for g in ref.values())
    worst = 0.0
    for k, p in net.named_parameters():
        g = ref[k].double()
        err = float((p.grad.cpu().double() - g).norm())
        assert err <= GRAD_TOL[gemm] * float(g.norm()) + (1e-6 if gemm == "simt" else 1e-5) * gscale, (label, k, err, float(g.norm()))
        if float(g.norm()) > 1e-3 * gscale:
            worst = max(worst, err / float(g.norm()))
    return worst


@pytest.mark.parametrize("gemm", GEMMS)
@pytest.mark.parametrize("name", ["train_wnet_a", "train_wnet_b"])
def test_loss_and_gradients_vs_reference(dwb, name, gemm):
    z, cfg = load(name)
    B, _, L = z["audio0"].shape
    net, tr = _trainer(dwb, cfg, sub(z, "sd0/"), B, L, gemm=gemm)
    loss, eps = tr.loss_backward(torch.from_numpy(z["audio0"]).cuda(), _dh(dwb, z), diffusion_steps=torch.from_numpy(z["steps0"]),
                                 z=torch.from_numpy(z["z0"]), return_eps=True)
    e = rel_l2(eps.cpu(), z["eps0"])
    worst = _check_grads(net, sub(z, "grad0/"), name, gemm)
    print(f"{name} [{gemm}]: loss {float(loss):.6f} (ref {z['losses'][0]:.6f}) eps rel_l2 {e:.2e} worst gradient rel_l2 {worst:.2e}")
    assert abs(float(loss) - z["losses"][0]) <= LOSS_TOL[gemm] * z["losses"][0] and e < LOSS_TOL[gemm]
    # the gradients ARE the module's .grad tensors (views of the flat buffer), and a second call overwrites them
    loss2 = tr.loss_backward(torch.from_numpy(z["audio0"]).cuda(), _dh(dwb, z), diffusion_steps=torch.from_numpy(z["steps0"]),
                             z=torch.from_numpy(z["z0"]))
    assert abs(float(loss2) - float(loss)) <= 1e-6 * float(loss)
    _check_grads(net, sub(z, "grad0/"), name + " (second call)", gemm)


@pytest.mark.parametrize("name", ["train_wnet_a", "train_wnet_b"])
def test_three_adam_steps_vs_reference(dwb, name):
    z, cfg = load(name)
    B, _, L = z["audio0"].shape
    net, tr = _trainer(dwb, cfg, sub(z, "sd0/"), B, L, lr=float(z["lr"]))
    dh = _dh(dwb, z)
    for it in range(3):
        loss = tr.loss_backward(torch.from_numpy(z[f"audio{it}"]).cuda(), dh, diffusion_steps=torch.from_numpy(z[f"steps{it}"]),
                                z=torch.from_numpy(z[f"z{it}"]))
        tr.step()
        assert abs(float(loss) - z["losses"][it]) <= 2e-5 * z["losses"][it], (it, float(loss), z["losses"][it])
    sd0, sd3, g0 = sub(z, "sd0/"), sub(z, "sd3/"), sub(z, "grad0/")
    has = {k[len("hasgrad/"):]: bool(z[k]) for k in z.files if k.startswith("hasgrad/")}
    gscale = max(float(g.norm()) for g in g0.values())
    after = {k: v.cpu() for k, v in net.state_dict().items()}
    worst = 0.0
    for k in sd0:
        if not has[k]:
            assert torch.equal(after[k], sd0[k]), k       # no gradient reaches it: Adam must leave it alone
            continue
        if float(g0[k].norm()) < 1e-6 * gscale:
            continue        # exact-zero gradient (d/dv of g v/|v| on a one-element row): Adam amplifies rounding noise
        moved = (sd3[k] - sd0[k]).double()
        err = float(((after[k] - sd0[k]).double() - moved).norm()) / float(moved.norm())
        worst = max(worst, err)
        assert err < 2e-2, (k, err)       # Adam's m/sqrt(v) turns 1e-6 gradient differences into up to ~1e-3 of a step
    print(f"{name}: worst relative difference of the 3-step parameter update {worst:.2e}")
    # torch.optim.Adam-format optimizer state round trip (checkpoints, train.py:158-160)
    osd = tr.state_dict()
    assert len(osd["state"]) == len(list(net.parameters())) and osd["param_groups"][0]["lr"] == float(z["lr"])
    m0 = tr.exp_avg.clone()
    tr.load_state_dict(osd)
    assert torch.equal(tr.exp_avg, m0) and tr.n_steps == 3


@pytest.mark.parametrize("gemm", GEMMS)
def test_gradients_vs_fp64_oracle_other_batch(dwb, gemm):
    """Seeded inputs that no fixture holds (ragged length, per-clip steps incl. 0 and T-1) against the fp64 oracle."""
    z, cfg = load("train_wnet_b")
    sd = sub(z, "sd0/")
    B, L = 2, 333
    g = torch.Generator().manual_seed(5)
    audio, zz = torch.rand(B, 1, L, generator=g) * 2 - 1, torch.randn(B, 1, L, generator=g)
    steps = torch.tensor([0, 49])
    net, tr = _trainer(dwb, cfg, sd, B, L, gemm=gemm)
    dh = _dh(dwb, z)
    loss = tr.loss_backward(audio.cuda(), dh, diffusion_steps=steps, z=zz)
    lo, _, go = TO.loss_and_grads_manual(cfg, sd, audio, steps, zz, dh["Alpha_bar"].cpu())
    assert abs(float(loss) - float(lo)) <= LOSS_TOL[gemm] * float(lo)
    worst = _check_grads(net, go, "oracle", gemm)
    print(f"fp64 oracle [{gemm}], B={B} L={L}: worst gradient rel_l2 {worst:.2e}")


@pytest.mark.parametrize("gemm", GEMMS)
def test_full_size_gradients_vs_reference(dwb, gemm):
    """BASELINE configs[0] (wnet h128/d30) at L = 16000: loss, the norm of every parameter gradient and six complete
    gradient tensors against the reference's autograd on the same seeded weights."""
    z, cfg = load("train_full_wnet_h128_d30")
    sd = dwb.init.seeded_state_dict(dict(cfg), seed=int(z["seed"]))
    g = torch.Generator().manual_seed(int(z["xseed"]))
    audio = torch.rand(1, 1, 16000, generator=g) * 2 - 1
    steps = torch.tensor([int(z["step"])])
    zz = torch.randn(1, 1, 16000, generator=g)
    net, tr = _trainer(dwb, cfg, sd, 1, 16000, gemm=gemm)
    dh = dwb.calc_diffusion_hyperparams(200, 1e-4, 0.02)
    loss = tr.loss_backward(audio.cuda(), dh, diffusion_steps=steps, z=zz)
    assert abs(float(loss) - float(z["loss"])) <= 2 * LOSS_TOL[gemm] * float(z["loss"]), (float(loss), float(z["loss"]))
    names, norms = [str(n) for n in z["names"]], z["grad_norms"]
    grads = {k: p.grad for k, p in net.named_parameters()}
    gscale = norms.max()
    worst = 0.0
    for k, n in zip(names, norms):
        mine = float(grads[k].double().norm())
        assert abs(mine - n) <= 1e-3 * n + 1e-6 * gscale, (k, mine, n)
        if n > 1e-3 * gscale:
            worst = max(worst, abs(mine - n) / n)
    full = 0.0
    for k in z.files:
        if k.startswith("grad/"):
            if float(np.linalg.norm(z[k])) < 1e-6 * gscale:
                continue    # exact-zero gradient (weight_v of the 1-input-channel init conv): fp32 rounding noise on both sides
            e = rel_l2(grads[k[5:]].cpu(), z[k])
            full = max(full, e)
            assert e < 1e-3, (k, e)
    info = tr.info()
    print(f"wnet h128/d30 L=16000 [{gemm}]: loss {float(loss):.6f}, worst gradient-norm difference {worst:.2e}, worst full-tensor rel_l2 {full:.2e}, "
          f"{info['launches']} launches, workspace {info['workspace_bytes'] / 2**20:.0f} MiB")


def test_training_reduces_the_loss_and_feeds_the_sampler(dwb):
    """Twenty Adam steps on one fixed batch lower its loss, and the trained parameters are what the inference engine
    then runs (the plan is rebuilt after step())."""
    z, cfg = load("train_wnet_a")
    B, _, L = z["audio0"].shape
    net, tr = _trainer(dwb, cfg, sub(z, "sd0/"), B, L, lr=2e-3)
    dh = _dh(dwb, z)
    audio, steps, zz = torch.from_numpy(z["audio0"]).cuda(), torch.from_numpy(z["steps0"]), torch.from_numpy(z["z0"])
    losses = []
    for _ in range(20):
        losses.append(float(tr.loss_backward(audio, dh, diffusion_steps=steps, z=zz)))
        tr.step()
    assert losses[-1] < 0.9 * losses[0], losses
    loss, eps = tr.loss_backward(audio, dh, diffusion_steps=steps, z=zz, return_eps=True)
    ab = dh["Alpha_bar"].cpu()[steps].view(B, 1, 1)
    x_t = (torch.sqrt(ab) * torch.from_numpy(z["audio0"]) + torch.sqrt(1 - ab) * zz).cuda()
    net.eval()
    with torch.no_grad():
        eps_inf = net((x_t, steps.view(B, 1).float().cuda()))
    assert rel_l2(eps_inf.cpu(), eps.cpu()) < 1e-4


def test_train_entry_point_checkpoints_and_resumes(dwb, tmp_path, monkeypatch, capsys):
    """train.py with the reference's override syntax: a few iterations on synthetic clips, a checkpoint in the
    reference's format, and a second invocation that resumes from it (train.py:94-114,155-160)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dwb_train_entry", os.path.join(root, "train.py"))
    entry = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(entry)
    monkeypatch.chdir(tmp_path)
    args = ["model=wavenet_small", "model.res_channels=16", "model.skip_channels=16", "model.num_res_layers=4", "model.dilation_cycle=2",
            "dataset.segment_length=512", "train.synthetic=true", "train.batch_size_per_gpu=2", "train.iters_per_ckpt=2",
            "train.iters_per_logging=1"]
    entry.main(args + ["train.n_iters=2"])
    out = capsys.readouterr().out
    assert "training from scratch" in out and "model at iteration 2 is saved" in out
    ckpts = [os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs if f.endswith(".pkl")]
    assert sorted(os.path.basename(c) for c in ckpts) == ["0.pkl", "2.pkl"]
    ck = torch.load([c for c in ckpts if c.endswith("2.pkl")][0], map_location="cpu")
    assert set(ck) == {"model_state_dict", "optimizer_state_dict"}
    net = dwb.construct_model(dict(_name_="wavenet", unconditional=True, res_channels=16, skip_channels=16, num_res_layers=4, dilation_cycle=2))
    net.load_state_dict(ck["model_state_dict"])                                   # the reference's keys
    opt = torch.optim.Adam(net.parameters(), lr=2e-4)
    opt.load_state_dict(ck["optimizer_state_dict"])                               # torch.optim.Adam reads our optimizer state
    assert float(ck["model_state_dict"]["final_conv.2.conv.weight"].abs().max()) > 0   # the zero-initialised head has moved
    entry.main(args + ["train.n_iters=3"])
    assert "Successfully loaded model at iteration 2" in capsys.readouterr().out
