"""Run-directory naming and checkpoint discovery, format-compatible with the reference
(utils.py:23-74, 96-116, 155-177; generate.py:98-121,188-192) so its exp/ trees can be used as is:
    exp/<model id>_T{T}_betaT{beta_T}[_L{seg}_hop{hop}]_{uncond|cond}/{checkpoint,waveforms/<iter>}/
    checkpoint file  <iter>.pkl = torch.save({'model_state_dict', 'optimizer_state_dict'})
    wav file         {iter//1000}k_{n_samples*rank+i}.wav, float32
"""
import os
import re

import torch

from .models import model_identifier


def run_id(name, model_cfg, diffusion_cfg, dataset_cfg):
    s = model_identifier(model_cfg) + f"_T{diffusion_cfg['T']}_betaT{diffusion_cfg['beta_T']}"
    if not model_cfg["unconditional"]:
        s += f"_L{dataset_cfg['segment_length']}_hop{dataset_cfg['hop_length']}"
    s += "_uncond" if model_cfg["unconditional"] else "_cond"
    return f"{name}_{s}" if name else s


def local_directory(name, model_cfg, diffusion_cfg, dataset_cfg, output_directory, root="exp"):
    local_path = run_id(name, model_cfg, diffusion_cfg, dataset_cfg)
    out = os.path.join(root, local_path, output_directory)
    os.makedirs(out, mode=0o775, exist_ok=True)
    return local_path, out


def checkpoint_iters(path):
    its = []
    for f in os.listdir(path):
        m = re.fullmatch(r"(\d+)\.pkl", f)
        if m:
            its.append(int(m.group(1)))
    return sorted(its)


def find_max_epoch(path):
    its = checkpoint_iters(path)
    return its[-1] if its else -1


def load_state_dict(ckpt_dir, ckpt_iter="max", ckpt_smooth=None):
    """-> (iteration, model_state_dict).  ckpt_smooth = arithmetic mean of every checkpoint in
    (ckpt_smooth, ckpt_iter] (the reference's experimental averaging, utils.py:47-74)."""
    it = find_max_epoch(ckpt_dir) if ckpt_iter == "max" else int(ckpt_iter)
    if ckpt_smooth is None:
        path = os.path.join(ckpt_dir, f"{it}.pkl")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"No valid model found: {path}")
        return it, torch.load(path, map_location="cpu")["model_state_dict"]
    chosen = [i for i in checkpoint_iters(ckpt_dir) if ckpt_smooth < i <= it]
    if not chosen:
        raise FileNotFoundError(f"no checkpoints in ({ckpt_smooth}, {it}] under {ckpt_dir}")
    avg = None
    for n, i in enumerate(chosen):
        sd = torch.load(os.path.join(ckpt_dir, f"{i}.pkl"), map_location="cpu")["model_state_dict"]
        avg = sd if avg is None else {k: (avg[k] * n + sd[k]) / (n + 1) for k in avg}
    return it, avg
