"""CPU oracle of the WaveNet TRAINING step - TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).

Restates, from the maths, what `train.py:137-143,198-222` does for `model._name_ = wavenet` (unconditional):
    x_t = sqrt(abar_t) x + sqrt(1 - abar_t) z;  loss = mean((net((x_t, t)) - z)^2);  loss.backward();  Adam.step()
in two independent forms:
  * `loss_and_grads_autograd`  torch autograd through the oracle's own forward (`diffwave_oracle.forward`; BOTH backbones - the
                               SaShiMi form is the pinned oracle for the half of the training row whose kernels are not built)
  * `loss_and_grads_manual`    the hand-derived backward written out layer by layer - the derivation the CUDA
                               kernels of csrc/train_wavenet.cu implement (same shift conventions and scale factors)
Pinned: tests/test_train_oracle.py checks both against tests/golden/train_wnet_*.npz, which
tests/golden/make_golden.py --train produced with the unmodified reference modules, torch autograd and
torch.optim.Adam.  Parity status: PINNED.
"""
import math

import torch

from . import diffwave_oracle as O


def training_loss(cfg, sd, audio, steps, z, alpha_bar, dtype=torch.float64):
    """train.py:198-222 with the draws passed in.  -> (loss, eps).  Either backbone (O.forward dispatches on cfg['_name_']):
    the SaShiMi form regenerates every S4 kernel from the parameters inside the graph, as the reference does per step
    (models/s4.py:674-807), so autograd reaches C, B, P, inv_w_real, w_imag and log_dt."""
    B = audio.shape[0]
    ab = alpha_bar[steps.reshape(B).long()].reshape(B, 1, 1)           # fp32 table lookups, like the reference
    x_t = (torch.sqrt(ab) * audio + torch.sqrt(1 - ab) * z) if dtype == torch.float32 else (
        torch.sqrt(ab).to(dtype) * audio.to(dtype) + torch.sqrt(1 - ab).to(dtype) * z.to(dtype))
    eps = O.forward(cfg, sd, x_t, steps.reshape(B, 1).to(dtype), None, dtype=dtype)
    return ((eps - z.to(dtype)) ** 2).mean(), eps


def loss_and_grads_autograd(cfg, sd, audio, steps, z, alpha_bar, dtype=torch.float64):
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full = {**sd, **leaves}          # integer buffers (SaShiMi's kernel.L) ride along untouched
    loss, eps = training_loss(cfg, full, audio, steps, z, alpha_bar, dtype)
    loss.backward()
    return loss.detach(), eps.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


def _shift(x, s):
    """y[..., l] = x[..., l + s] inside [0, L), 0 outside."""
    L = x.shape[-1]
    y = torch.zeros_like(x)
    if s >= 0:
        y[..., :L - s] = x[..., s:]
    else:
        y[..., -s:] = x[..., :L + s]
    return y


def _wn_fold(v, g):
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
    return v * (g / n)


def _wn_unfold(v, g, dW):
    """dg, dv of W = g v / |v| given dW."""
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
    dot = (dW * v).reshape(v.shape[0], -1).sum(1).reshape(-1, 1, 1)
    return dot / n, (g / n) * dW - (g * dot / n ** 3) * v


def _dswish(z):
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def loss_and_grads_manual(cfg, sd, audio, steps, z, alpha_bar, dtype=torch.float64):
    """The training step's backward written out by hand (no autograd anywhere).  -> (loss, eps, {key: grad})"""
    with torch.no_grad():
        sd = {k: v.to(dtype) for k, v in sd.items() if v.is_floating_point()}
        C, N, cycle = cfg["res_channels"], cfg["num_res_layers"], cfg["dilation_cycle"]
        B, _, L = audio.shape
        rs2, cN = math.sqrt(0.5), math.sqrt(1.0 / N)
        ab = alpha_bar[steps.reshape(B).long()].reshape(B, 1, 1)
        xt = torch.sqrt(ab).to(dtype) * audio.to(dtype) + torch.sqrt(1 - ab).to(dtype) * z.to(dtype)
        z = z.to(dtype)
        G = {}
        # ---- forward, keeping what the backward needs ----
        e0 = O.step_embedding(steps.reshape(B, 1).to(dtype), cfg.get("diffusion_step_embed_dim_in", 128))
        W1, b1, W2, b2 = (sd["residual_layer.fc_t1.weight"], sd["residual_layer.fc_t1.bias"], sd["residual_layer.fc_t2.weight"],
                          sd["residual_layer.fc_t2.bias"])
        z1 = e0 @ W1.T + b1
        e1 = z1 * torch.sigmoid(z1)
        z2 = e1 @ W2.T + b2
        e2 = z2 * torch.sigmoid(z2)
        p0 = "init_conv.0.conv"
        Win = _wn_fold(sd[p0 + ".weight_v"], sd[p0 + ".weight_g"])
        H = [torch.relu(torch.einsum("mk,bkl->bml", Win[:, :, 0], xt) + sd[p0 + ".bias"][None, :, None])]
        TH, SG, Os, parts, eff = [], [], [], [], []
        SK = torch.zeros(B, cfg["skip_channels"], L, dtype=dtype)
        for n in range(N):
            p = f"residual_layer.residual_blocks.{n}."
            d = 2 ** (n % cycle)
            part = e2 @ sd[p + "fc_t.weight"].T + sd[p + "fc_t.bias"]
            Wd = _wn_fold(sd[p + "dilated_conv_layer.conv.weight_v"], sd[p + "dilated_conv_layer.conv.weight_g"])
            Wr = _wn_fold(sd[p + "res_conv.weight_v"], sd[p + "res_conv.weight_g"])
            Ws = _wn_fold(sd[p + "skip_conv.weight_v"], sd[p + "skip_conv.weight_g"])
            u = H[n] + part[:, :, None]
            g = sd[p + "dilated_conv_layer.conv.bias"][None, :, None] + sum(
                torch.einsum("mk,bkl->bml", Wd[:, :, k], _shift(u, (k - 1) * d)) for k in range(3))
            th, sg = torch.tanh(g[:, :C]), torch.sigmoid(g[:, C:])
            o = th * sg
            H.append((H[n] + torch.einsum("mk,bkl->bml", Wr[:, :, 0], o) + sd[p + "res_conv.bias"][None, :, None]) * rs2)
            SK = SK + torch.einsum("mk,bkl->bml", Ws[:, :, 0], o) + sd[p + "skip_conv.bias"][None, :, None]
            TH.append(th); SG.append(sg); Os.append(o); parts.append(part); eff.append((Wd, Wr, Ws))
        pf = "final_conv.0.conv"
        Wf = _wn_fold(sd[pf + ".weight_v"], sd[pf + ".weight_g"])
        Fh = torch.relu(cN * torch.einsum("mk,bkl->bml", Wf[:, :, 0], SK) + sd[pf + ".bias"][None, :, None])
        Wz, bz = sd["final_conv.2.conv.weight"], sd["final_conv.2.conv.bias"]
        y = torch.einsum("mk,bkl->bml", Wz[:, :, 0], Fh) + bz[None, :, None]
        loss = ((y - z) ** 2).mean()
        # ---- backward: head ----
        dY = 2 * (y - z) / y.numel()
        G["final_conv.2.conv.weight"] = torch.einsum("bml,bkl->mk", dY, Fh)[:, :, None]
        G["final_conv.2.conv.bias"] = dY.sum((0, 2))
        dF = torch.einsum("mk,bml->bkl", Wz[:, :, 0], dY) * (Fh > 0)
        dWf = cN * torch.einsum("bml,bkl->mk", dF, SK)[:, :, None]
        G[pf + ".bias"] = dF.sum((0, 2))
        G[pf + ".weight_g"], G[pf + ".weight_v"] = _wn_unfold(sd[pf + ".weight_v"], sd[pf + ".weight_g"], dWf)
        dS = cN * torch.einsum("mk,bml->bkl", Wf[:, :, 0], dF)
        # ---- backward: residual layers ----
        dH = torch.zeros_like(H[0])
        de2 = torch.zeros_like(e2)
        for n in range(N - 1, -1, -1):
            p = f"residual_layer.residual_blocks.{n}."
            d = 2 ** (n % cycle)
            Wd, Wr, Ws = eff[n]
            th, sg, o = TH[n], SG[n], Os[n]
            dO = rs2 * torch.einsum("mk,bml->bkl", Wr[:, :, 0], dH) + torch.einsum("mk,bml->bkl", Ws[:, :, 0], dS)
            dWr = rs2 * torch.einsum("bml,bkl->mk", dH, o)[:, :, None]
            dWs = torch.einsum("bml,bkl->mk", dS, o)[:, :, None]
            G[p + "res_conv.bias"] = rs2 * dH.sum((0, 2))
            G[p + "skip_conv.bias"] = dS.sum((0, 2))
            G[p + "res_conv.weight_g"], G[p + "res_conv.weight_v"] = _wn_unfold(sd[p + "res_conv.weight_v"], sd[p + "res_conv.weight_g"], dWr)
            G[p + "skip_conv.weight_g"], G[p + "skip_conv.weight_v"] = _wn_unfold(sd[p + "skip_conv.weight_v"], sd[p + "skip_conv.weight_g"], dWs)
            dG = torch.cat([dO * sg * (1 - th * th), dO * th * sg * (1 - sg)], 1)
            dU = sum(torch.einsum("mk,bml->bkl", Wd[:, :, k], _shift(dG, -(k - 1) * d)) for k in range(3))
            u = H[n] + parts[n][:, :, None]
            dWd = torch.stack([torch.einsum("bml,bkl->mk", dG, _shift(u, (k - 1) * d)) for k in range(3)], 2)
            pd = p + "dilated_conv_layer.conv"
            G[pd + ".bias"] = dG.sum((0, 2))
            G[pd + ".weight_g"], G[pd + ".weight_v"] = _wn_unfold(sd[pd + ".weight_v"], sd[pd + ".weight_g"], dWd)
            dpart = dU.sum(2)
            G[p + "fc_t.weight"] = dpart.T @ e2
            G[p + "fc_t.bias"] = dpart.sum(0)
            de2 = de2 + dpart @ sd[p + "fc_t.weight"]
            dH = dU + rs2 * dH
        # ---- backward: input conv, step embedding ----
        dpre = dH * (H[0] > 0)
        dWin = torch.einsum("bml,bkl->mk", dpre, xt)[:, :, None]
        G[p0 + ".bias"] = dpre.sum((0, 2))
        G[p0 + ".weight_g"], G[p0 + ".weight_v"] = _wn_unfold(sd[p0 + ".weight_v"], sd[p0 + ".weight_g"], dWin)
        dz2 = de2 * _dswish(z2)
        G["residual_layer.fc_t2.weight"], G["residual_layer.fc_t2.bias"] = dz2.T @ e1, dz2.sum(0)
        dz1 = (dz2 @ W2) * _dswish(z1)
        G["residual_layer.fc_t1.weight"], G["residual_layer.fc_t1.bias"] = dz1.T @ e0, dz1.sum(0)
        return loss, y, G


def adam_step(p, g, m, v, lr, step, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam (amsgrad off, no weight decay), in place on tensors of any dtype; `step` counts from 1."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
    p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
