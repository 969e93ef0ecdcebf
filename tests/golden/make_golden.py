"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (CPU, fp32).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [--full]

Everything here goes through oracle/refshim.py (SURVEY.md Appendix C recipe: 4 module stubs,
`.cuda()` -> identity, conjugate-symmetric Cauchy fallback).  Outputs are small .npz files in
this directory; `-m gpu` tests and the oracle tests read them, never /root/reference.

Fixture families
  schedule_*.npz      utils.calc_diffusion_hyperparams                      (utils.py:121-151)
  embed.npz           models.utils.calc_diffusion_step_embedding            (models/utils.py:4-29)
  cauchy.npz          extensions/cauchy/cauchy.py:cauchy_mult_torch in complex128 on the inputs of
                      extensions/cauchy/test_cauchy.py:11-23,53-66 (seed 2357, batch 4)
  s4kernel_*.npz      models.s4.S4(...).kernel(L) before/after _setup_C      (s4.py:525-551,674-807)
  tiny_*.npz          full state_dict + (x, t, mel) -> eps for tiny WaveNet / SaShiMi models
  traj_*.npz          generate.sampling() on a tiny model, seeded            (generate.py:23-55)
  full_*.npz (--full) eps of the reference at BASELINE sizes for weights created by OUR
                      package's seeded initialiser (only the 64 KB outputs are stored)
"""
import argparse
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

SMALL_EMB = dict(diffusion_step_embed_dim_in=16, diffusion_step_embed_dim_mid=32,
                 diffusion_step_embed_dim_out=32)
# Sashimi's DiffWaveBlock ignores the config and hard-wires fc_t to 512 inputs (sashimi.py:116)
SMALL_EMB_S = dict(diffusion_step_embed_dim_in=16, diffusion_step_embed_dim_mid=32,
                   diffusion_step_embed_dim_out=512)

TINY = {
    "tiny_unet": dict(base="unet_d64", over=dict(d_model=8, n_layers=2, L=256, **SMALL_EMB_S), B=2, L=256),
    "tiny_snet": dict(base="unet_d64", over=dict(d_model=8, n_layers=2, L=320, unet=False, pool=[4, 2], **SMALL_EMB_S), B=1, L=320),
    "tiny_unet_e128": dict(base="unet_d64", over=dict(d_model=4, n_layers=1, L=64, pool=[2, 2]), B=3, L=64),
    "tiny_unet_cond": dict(base="unet_d32_cond", over=dict(d_model=8, n_layers=1, L=512, **SMALL_EMB_S), B=2, L=512, mel=(1, 80, 3)),
    "tiny_unet_condB": dict(base="unet_d32_cond", over=dict(d_model=8, n_layers=1, L=512, **SMALL_EMB_S), B=2, L=512, mel=(2, 80, 2)),
    "tiny_wnet": dict(base="wnet_h128_d30", over=dict(res_channels=16, skip_channels=8, num_res_layers=7, dilation_cycle=5, **SMALL_EMB), B=2, L=300),
    "tiny_wnet_cond": dict(base="wnet_h128_d30", over=dict(res_channels=8, skip_channels=16, num_res_layers=3, dilation_cycle=2, unconditional=False, mel_upsample=[16, 16], **SMALL_EMB), B=2, L=512, mel=(1, 80, 2)),
}


def np_sd(sd):
    return {"sd/" + k: v.detach().cpu().numpy() for k, v in sd.items()}


def nonzero_final(net):
    """ZeroConv1d makes a fresh model output exactly 0 (SURVEY finding 5)."""
    with torch.no_grad():
        w = net.final_conv[2].conv.weight
        w.normal_(0, (1.0 / w.shape[1]) ** 0.5)
        net.final_conv[2].conv.bias.fill_(0.05)


def build(ns, base, over, seed=0):
    cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
    cfg.update(over)
    torch.manual_seed(seed)
    net = ns.models.construct_model(cfg).eval()
    nonzero_final(net)
    return cfg, net


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--train", action="store_true", help="training-step fixtures only (train_*.npz)")
    ap.add_argument("--extra", action="store_true", help="round-2 fixtures only (new files; the round-1 files are left untouched)")
    args = ap.parse_args()
    ns = refshim.load(parity=True)
    if args.extra:
        return make_extra(ns)
    if args.train:
        return make_train(ns)

    # --- schedule ---------------------------------------------------------------------
    for tag, kw in (("T200", dict(T=200, beta_0=1e-4, beta_T=0.02)), ("T50", dict(T=50, beta_0=1e-4, beta_T=0.05)),
                    ("fast6", dict(T=6, beta_0=1e-4, beta_T=0.02, beta=[0.0001, 0.001, 0.01, 0.05, 0.2, 0.5], fast=True))):
        dh = ns.utils.calc_diffusion_hyperparams(**kw)
        save("schedule_" + tag, T=dh["T"], **{k: dh[k].numpy() for k in ("Beta", "Alpha", "Alpha_bar", "Sigma")},
             kw=np.array(repr(kw)))

    # --- step embedding ---------------------------------------------------------------
    import models.utils as ref_mu  # imported by refshim under the reference's package name
    t = torch.tensor([[0.], [1.], [2.5], [49.], [100.], [199.]])
    save("embed", t=t.numpy(), e128=ref_mu.calc_diffusion_step_embedding(t, 128).numpy(),
         e16=ref_mu.calc_diffusion_step_embedding(t, 16).numpy())

    # --- Cauchy op (reference's own test recipe) --------------------------------------
    sys.modules["cauchy_mult"] = types.SimpleNamespace(cauchy_mult_fwd=None, cauchy_mult_bwd=None,
                                                       cauchy_mult_sym_fwd=None, cauchy_mult_sym_bwd=None)
    sys.path.insert(0, os.path.join(refshim.REF_ROOT, "extensions", "cauchy"))
    import cauchy as ref_cauchy
    sys.path.pop(0)
    out = {}
    for N in (4, 8, 64, 256):            # full-spectrum sizes as in test_cauchy.py:54 (kernel sees N/2)
        for L in (3, 17, 489, 1024, 1047):
            torch.random.manual_seed(2357)
            v_half = torch.randn(4, N // 2, dtype=torch.complex64)
            v = torch.cat([v_half, v_half.conj()], dim=-1)
            w_half = torch.randn(4, N // 2, dtype=torch.complex64)
            w = torch.cat([w_half, w_half.conj()], dim=-1)
            z = torch.exp(1j * torch.randn(L, dtype=torch.float32))
            ref = ref_cauchy.cauchy_mult_torch(v.cdouble(), z.cdouble(), w.cdouble(), symmetric=True)
            out[f"v_{N}_{L}"], out[f"w_{N}_{L}"], out[f"z_{N}_{L}"] = v_half.numpy(), w_half.numpy(), z.numpy()
            out[f"out_{N}_{L}"] = ref.numpy()
    save("cauchy", **out)

    # --- S4 kernel generation ---------------------------------------------------------
    for H, L in ((4, 64), (3, 100), (2, 250), (2, 1000)):
        torch.manual_seed(3)
        s4 = ns.s4.S4(H, l_max=L, bidirectional=True).eval()
        sd0 = {k: v.clone() for k, v in s4.state_dict().items()}
        with torch.no_grad():
            k, _ = s4.kernel(L=L, rate=1.0)
        sd1 = s4.state_dict()
        omega = s4.kernel.kernel.omega
        save(f"s4kernel_H{H}_L{L}", k=k.numpy(), omega=omega.numpy(),
             **{"sd0/" + kk: v.numpy() for kk, v in sd0.items()}, **{"sd1/" + kk: v.numpy() for kk, v in sd1.items()})

    # --- tiny models: weights + io ----------------------------------------------------
    for name, spec in TINY.items():
        cfg, net = build(ns, spec["base"], spec["over"])
        B, L = spec["B"], spec["L"]
        g = torch.Generator().manual_seed(11)
        x = torch.randn(B, 1, L, generator=g)
        t = torch.tensor([[3.], [17.], [0.]])[:B]
        mel = torch.randn(*spec["mel"], generator=g) if "mel" in spec else None
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        with torch.no_grad():
            eps = net((x, t), mel_spec=mel)
        arrs = dict(cfg=np.array(repr(dict(cfg))), x=x.numpy(), t=t.numpy(), eps=eps.numpy(), **np_sd(net.state_dict()))
        if mel is not None:
            arrs["mel"] = mel.numpy()
        if cfg["_name_"] == "sashimi":   # also keep the fresh (kernel.L == 0) C tensors to pin _setup_C
            arrs.update({"sd0/" + k: v.numpy() for k, v in sd0.items() if k.endswith("kernel.kernel.C") or k.endswith("kernel.kernel.L")})
        save(name, **arrs)

    # --- sampler trajectories on tiny models -----------------------------------------
    for name in ("tiny_unet", "tiny_wnet"):
        spec = TINY[name]
        cfg, net = build(ns, spec["base"], spec["over"])
        with torch.no_grad():
            net.final_conv[2].conv.weight.mul_(8.0)  # make eps matter against the injected noise
        B, L, T = 2, spec["L"], 12
        dh = ns.utils.calc_diffusion_hyperparams(T=T, beta_0=1e-4, beta_T=0.02, fast=True)
        with torch.no_grad():
            net((torch.zeros(1, 1, L), torch.zeros(1, 1)))   # settle _setup_C before saving weights
        torch.manual_seed(1234)
        x0 = ns.generate.sampling(net, (B, 1, L), dh)
        save("traj_" + name, cfg=np.array(repr(dict(cfg))), T=T, beta_0=1e-4, beta_T=0.02, seed=1234,
             x0=x0.numpy(), **np_sd(net.state_dict()))

    if args.full:
        make_full(ns)


FULL = {
    # name: (base cfg, B, t, mel)
    "full_wnet_h128_d30": ("wnet_h128_d30", 1, 100.0, None),
    "full_unet_d64": ("unet_d64", 1, 100.0, None),
    "full_unet_d32_cond": ("unet_d32_cond", 1, 25.0, (1, 80, 63)),
    "full_unet_d128": ("unet_d128", 1, 100.0, None),            # BASELINE.json configs[2] (widths 128/256/512)
    "full_wnet_h256_d36": ("wnet_h256_d36", 1, 100.0, None),    # BASELINE.json configs[4]
}


def make_full(ns):
    """Reference eps at BASELINE sizes for weights made by diffwave_sashimi_b200's own seeded
    initialiser: the GPU-box tests rebuild the same weights from the seed and compare the CUDA
    engine with these stored reference outputs (64 KB each)."""
    import diffwave_sashimi_b200 as dwb
    for name, (base, B, tval, melshape) in FULL.items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and os.environ.get("GOLDEN_ONLY_MISSING"):
            continue
        cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
        sd = dwb.init.seeded_state_dict(dict(cfg), seed=0)
        net = ns.models.construct_model(cfg).eval()
        net.load_state_dict(sd)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, 1, 16000, generator=g)
        mel = torch.randn(*melshape, generator=g) if melshape else None
        outs = {}
        with torch.no_grad():
            for tv in (tval, 0.0):
                outs[f"eps_t{int(tv)}"] = net((x, tv * torch.ones(B, 1)), mel_spec=mel).numpy()
        save(name, cfg=np.array(repr(dict(cfg))), seed=0, xseed=5, **outs)


TRAIN_TINY = {
    "train_wnet_a": dict(over=dict(res_channels=16, skip_channels=8, num_res_layers=7, dilation_cycle=5, **SMALL_EMB), B=2, L=300),
    # widths and a length that are not multiples of the 64-wide GEMM tiles, dilation up to 8 on 200 samples
    "train_wnet_b": dict(over=dict(res_channels=40, skip_channels=24, num_res_layers=5, dilation_cycle=4, **SMALL_EMB), B=3, L=200),
}


def ref_training_loss(net, audio, dh, steps, z):
    """train.py:198-222 with the two random draws (diffusion_steps, z) passed in instead of drawn."""
    B = audio.shape[0]
    ab = dh["Alpha_bar"]
    steps = steps.view(B, 1, 1)
    x_t = torch.sqrt(ab[steps]) * audio + torch.sqrt(1 - ab[steps]) * z
    eps = net((x_t, steps.view(B, 1),), mel_spec=None)
    return torch.nn.MSELoss()(eps, z), eps


def make_train(ns):
    """Training-step fixtures from the unmodified reference modules + torch autograd + torch.optim.Adam (train.py:84-143):
      train_wnet_{a,b}.npz     tiny WaveNets: start state_dict, three (audio, steps, z) batches, loss / eps / every
                               parameter gradient of batch 0, losses and the state_dict after three Adam steps (lr 2e-4)
      train_unet_tiny.npz      tiny SaShiMi UNet: loss, eps and every parameter gradient (S4 parameters included) of one batch
      train_full_wnet_h128_d30 BASELINE configs[0] at B=1, L=16000 on our seeded weights: loss, the L2 norm of every
                               parameter gradient and a few complete gradient tensors
    """
    import diffwave_sashimi_b200 as dwb
    dh = ns.utils.calc_diffusion_hyperparams(T=50, beta_0=1e-4, beta_T=0.05, fast=False)
    for name, spec in TRAIN_TINY.items():
        cfg, net = build(ns, "wnet_h128_d30", spec["over"])
        net.train()
        B, L = spec["B"], spec["L"]
        g = torch.Generator().manual_seed(21)
        arrs = dict(cfg=np.array(repr(dict(cfg))), T=50, beta_0=1e-4, beta_T=0.05, lr=2e-4,
                    **{"sd0/" + k: v.detach().clone().numpy() for k, v in net.state_dict().items()})
        opt = torch.optim.Adam(net.parameters(), lr=2e-4)
        losses = []
        for it in range(3):
            audio = torch.rand(B, 1, L, generator=g) * 2 - 1
            steps = torch.randint(50, size=(B,), generator=g)
            z = torch.randn(B, 1, L, generator=g)
            opt.zero_grad()
            loss, eps = ref_training_loss(net, audio, dh, steps, z)
            loss.backward()
            if it == 0:
                arrs["eps0"] = eps.detach().numpy()
                for k, p in net.named_parameters():
                    arrs["grad0/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone().numpy()
                    arrs["hasgrad/" + k] = np.array(p.grad is not None)
            opt.step()
            losses.append(float(loss))
            arrs[f"audio{it}"], arrs[f"steps{it}"], arrs[f"z{it}"] = audio.numpy(), steps.numpy(), z.numpy()
        arrs["losses"] = np.array(losses, dtype=np.float64)
        arrs.update({"sd3/" + k: v.detach().clone().numpy() for k, v in net.state_dict().items()})
        save(name, **arrs)

    # SaShiMi (oracle pin for the half of the row without kernels): the tiny UNet of tiny_unet.npz with settled kernels
    spec = TINY["tiny_unet"]
    cfg, net = build(ns, spec["base"], spec["over"])
    B, L = spec["B"], spec["L"]
    g = torch.Generator().manual_seed(23)
    audio = torch.rand(B, 1, L, generator=g) * 2 - 1
    steps = torch.randint(50, size=(B,), generator=g)
    z = torch.randn(B, 1, L, generator=g)
    with torch.no_grad():
        net((audio, steps.view(B, 1).float()))            # _setup_C (models/s4.py:525-551) before the weights are saved
    net.train()
    settled = np.load(os.path.join(HERE, "tiny_unet.npz"))         # the settled weights are input independent: stored once, there
    assert all(np.array_equal(settled["sd/" + k], v.detach().numpy()) for k, v in net.state_dict().items())
    loss, eps = ref_training_loss(net, audio, dh, steps, z)
    loss.backward()
    save("train_unet_tiny", cfg=np.array(repr(dict(cfg))), T=50, beta_0=1e-4, beta_T=0.05, audio0=audio.numpy(), steps0=steps.numpy(),
         z0=z.numpy(), losses=np.array([float(loss)]), eps0=eps.detach().numpy(), weights=np.array("tiny_unet.npz sd/"),
         **{"grad0/" + k: p.grad.numpy() for k, p in net.named_parameters()})

    base = "wnet_h128_d30"
    cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
    net = ns.models.construct_model(cfg).train()
    net.load_state_dict(dwb.init.seeded_state_dict(dict(cfg), seed=0))
    dh = ns.utils.calc_diffusion_hyperparams(T=200, beta_0=1e-4, beta_T=0.02, fast=False)
    g = torch.Generator().manual_seed(22)
    audio = torch.rand(1, 1, 16000, generator=g) * 2 - 1
    steps = torch.tensor([137])
    z = torch.randn(1, 1, 16000, generator=g)
    loss, _ = ref_training_loss(net, audio, dh, steps, z)
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    norms = np.array([float(p.grad.double().norm()) if p.grad is not None else 0.0 for _, p in net.named_parameters()])
    keep = ["residual_layer.fc_t2.bias", "final_conv.2.conv.weight", "residual_layer.residual_blocks.0.dilated_conv_layer.conv.bias",
            "residual_layer.residual_blocks.29.skip_conv.weight_g", "residual_layer.residual_blocks.13.res_conv.weight_v",
            "init_conv.0.conv.weight_v"]
    gd = dict(net.named_parameters())
    save("train_full_" + base, cfg=np.array(repr(dict(cfg))), seed=0, xseed=22, step=137, loss=float(loss), names=np.array(names),
         grad_norms=norms, **{"grad/" + k: gd[k].grad.numpy() for k in keep})


def make_extra(ns):
    """Round-2 fixtures (all from the unmodified reference):
      full2_<cfg>.npz            eps at B=2 with DISTINCT steps per row, t = (1, T-1), all five BASELINE configs
      trajfull_<cfg>.npz         x_0 of the complete T-step generate.sampling() at BASELINE size with the last conv
                                 rescaled so that std(eps) ~ 1 (otherwise x_0 is all injected noise)
      lengths_<tiny>.npz         eps of tiny models at sequence lengths below and above the configured l_max
                                 (kernel truncation / L > l_max, models/s4.py:1387,1403-1406)
      full_unet_d32_cond_L51200  the LJSpeech vocoder config on a 200-frame utterance (generate.py:135-156)
      s4double_*.npz             kernel-length doubling of a set-up kernel (models/s4.py:531-534)
    """
    import diffwave_sashimi_b200 as dwb
    only = os.environ.get("GOLDEN_EXTRA", "").split(",") if os.environ.get("GOLDEN_EXTRA") else None
    want = lambda tag: only is None or tag in only

    if want("double"):
        for H, L in ((2, 250), (3, 64)):
            torch.manual_seed(3)
            s4 = ns.s4.S4(H, l_max=L, bidirectional=True).eval()
            with torch.no_grad():
                s4.kernel(L=L, rate=1.0)                   # _setup_C at L
                sd1 = {k: v.clone() for k, v in s4.state_dict().items()}
                k2, _ = s4.kernel(L=2 * L, rate=1.0)       # doubles to 2L (s4.py:723-724 -> :531-534)
                sd2 = {k: v.clone() for k, v in s4.state_dict().items()}
                k4, _ = s4.kernel(L=4 * L, rate=1.0)
                sd4 = {k: v.clone() for k, v in s4.state_dict().items()}
            save(f"s4double_H{H}_L{L}", k2=k2.numpy(), k4=k4.numpy(),
                 **{"sd0/" + kk: v.numpy() for kk, v in sd1.items()}, **{"sd1/" + kk: v.numpy() for kk, v in sd2.items()},
                 **{"sd/" + kk: v.numpy() for kk, v in sd4.items()})

    if want("lengths"):
        for name, lens, hop in (("tiny_unet", (128, 64, 512, 1024, 2304), None), ("tiny_snet", (160, 640), None),
                                ("tiny_unet_cond", (256, 1024, 2048), 256)):
            spec = TINY[name]
            cfg, net = build(ns, spec["base"], spec["over"])
            with torch.no_grad():                           # same settled weights as <name>.npz
                g = torch.Generator().manual_seed(11)
                x = torch.randn(spec["B"], 1, spec["L"], generator=g)
                t = torch.tensor([[3.], [17.], [0.]])[:spec["B"]]
                mel = torch.randn(*spec["mel"], generator=g) if "mel" in spec else None
                net((x, t), mel_spec=mel)
            arrs = {}
            for Lr in lens:
                g = torch.Generator().manual_seed(100 + Lr)
                x = torch.randn(2, 1, Lr, generator=g)
                t = torch.tensor([[5.], [40.]])
                mel = torch.randn(1, 80, Lr // hop + 1, generator=g) if hop else None
                with torch.no_grad():
                    eps = net((x, t), mel_spec=mel)
                arrs[f"x_{Lr}"], arrs[f"eps_{Lr}"] = x.numpy(), eps.numpy()
                if mel is not None:
                    arrs[f"mel_{Lr}"] = mel.numpy()
            save("lengths_" + name, t=t.numpy(), **arrs)

    if want("mel"):
        # dataloaders/stft.py imports librosa (absent here) for two helpers: pad_center (a no-op at win_length ==
        # filter_length) and filters.mel, for which torchaudio's independent implementation of the same Slaney
        # filterbank stands in.  Everything else (windowed Fourier basis, reflect padding, conv1d, magnitude, log) is
        # the reference's own code.
        import torchaudio
        lb = types.ModuleType("librosa")
        lb.util = types.SimpleNamespace(pad_center=lambda w, size: np.pad(w, ((size - len(w)) // 2, size - len(w) - (size - len(w)) // 2)),
                                        tiny=lambda x: np.finfo(np.float32).tiny)
        lb.filters = types.SimpleNamespace(mel=lambda sr, n_fft, n_mels, fmin, fmax: torchaudio.functional.melscale_fbanks(
            n_fft // 2 + 1, fmin, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy())
        sys.modules["librosa"] = lb
        import importlib.util           # the package __init__ pulls in dataset downloaders; load the one file
        spec = importlib.util.spec_from_file_location("ref_stft", os.path.join(refshim.REF_ROOT, "dataloaders", "stft.py"))
        ref_stft = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_stft)
        arrs = {}
        for tag, kw, T in (("lj", dict(filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80,
                                        sampling_rate=22050, mel_fmin=0.0, mel_fmax=8000.0), 16000),
                           ("small", dict(filter_length=256, hop_length=64, win_length=200, n_mel_channels=20,
                                          sampling_rate=16000, mel_fmin=50.0, mel_fmax=7000.0), 3001)):
            st = ref_stft.TacotronSTFT(**kw)
            g = torch.Generator().manual_seed(31)
            wav = (torch.rand(2, T, generator=g) * 2 - 1) * torch.tensor([[0.9], [0.05]])     # a loud and a quiet clip
            wav[1, T // 2:] = 0.0                                                               # silence -> the clamp
            with torch.no_grad():
                m = st.mel_spectrogram(wav)
            arrs[f"wav_{tag}"], arrs[f"mel_{tag}"], arrs[f"kw_{tag}"] = wav.numpy(), m.numpy(), np.array(repr(kw))
            arrs[f"basis_{tag}"] = st.stft_fn.forward_basis[:, 0, :].numpy()[:, ::max(1, kw["filter_length"] // 64)]
        save("mel_frontend", **arrs)

    T_OF = {"wnet_h128_d30": 200, "unet_d64": 200, "unet_d32_cond": 50, "unet_d128": 200, "wnet_h256_d36": 200}
    if want("full2"):
        for name, (base, _, _, melshape) in FULL.items():
            T = T_OF[base]
            cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
            net = ns.models.construct_model(cfg).eval()
            net.load_state_dict(dwb.init.seeded_state_dict(dict(cfg), seed=0))
            g = torch.Generator().manual_seed(6)
            x = torch.randn(2, 1, 16000, generator=g)
            mel = torch.randn(*melshape, generator=g) if melshape else None
            t = torch.tensor([[1.0], [float(T - 1)]])
            with torch.no_grad():
                eps = net((x, t), mel_spec=mel).numpy()
            save(name.replace("full_", "full2_"), cfg=np.array(repr(dict(cfg))), seed=0, xseed=6, t=t.numpy(), eps=eps)

    if want("long"):
        base, frames = "unet_d32_cond", 200
        cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
        net = ns.models.construct_model(cfg).eval()
        net.load_state_dict(dwb.init.seeded_state_dict(dict(cfg), seed=0))
        g = torch.Generator().manual_seed(7)
        Lr = frames * 256
        x = torch.randn(1, 1, Lr, generator=g)
        mel = torch.randn(1, 80, frames, generator=g)
        with torch.no_grad():
            eps = net((x, torch.full((1, 1), 25.0)), mel_spec=mel).numpy()
        save("full_unet_d32_cond_L51200", cfg=np.array(repr(dict(cfg))), seed=0, xseed=7, frames=frames, t=25.0, eps=eps)

    if want("traj"):
        for base, melshape, beta_T in (("unet_d32_cond", (1, 80, 63), 0.05), ("unet_d64", None, 0.02)):
            T = T_OF[base]
            cfg = refshim.Cfg(refshim.MODEL_CFGS[base])
            net = ns.models.construct_model(cfg).eval()
            net.load_state_dict(dwb.init.seeded_state_dict(dict(cfg), seed=0))
            g = torch.Generator().manual_seed(8)
            mel = torch.randn(*melshape, generator=g) if melshape else None
            x = torch.randn(1, 1, 16000, generator=g)
            with torch.no_grad():
                e = net((x, torch.full((1, 1), float(T // 2))), mel_spec=mel)
                scale = float(1.0 / e.std())
                net.final_conv[2].conv.weight.mul_(scale)
                net.final_conv[2].conv.bias.mul_(scale)
                e2 = net((x, torch.full((1, 1), float(T // 2))), mel_spec=mel)
            dh = ns.utils.calc_diffusion_hyperparams(T=T, beta_0=1e-4, beta_T=beta_T, fast=True)
            torch.manual_seed(4242)
            x0 = ns.generate.sampling(net, (1, 1, 16000), dh, condition=mel)
            save("trajfull_" + base, cfg=np.array(repr(dict(cfg))), seed=0, melseed=8, noise_seed=4242, T=T, beta_0=1e-4,
                 beta_T=beta_T, final_scale=scale, eps_std_after=float(e2.std()), x0=x0.numpy())


if __name__ == "__main__":
    main()
