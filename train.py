#!/usr/bin/env python
"""Training entry point with the reference's command line (train.py:224-250):

    python train.py experiment=sc09 model=wavenet_small train.batch_size_per_gpu=4
    python -m torch.distributed.run --nproc-per-node 8 train.py model=wavenet ...     # one process per GPU

The loop is the reference's (train.py:84-195): load the newest checkpoint if there is one, then per batch
zero_grad / training_loss / backward / (gradient all-reduce) / Adam step, a `<iter>.pkl` checkpoint every
`train.iters_per_ckpt` iterations in the reference's format ({'model_state_dict', 'optimizer_state_dict'}, readable by
the reference and by generate.py).  What runs underneath differs: forward, backward and Adam are libdwb kernels
(diffwave_sashimi_b200.training.Trainer; no autograd), the gradients of all ranks are ONE contiguous buffer reduced by
one NCCL all-reduce per step, and ranks start from rank 0's parameters by one broadcast.

Built for model._name_ = wavenet, unconditional.  SaShiMi / mel-conditioned training raise (their backward kernels are
not written).  Data: every .wav under dataset.data_path (zero-padded / cropped to dataset.segment_length, int16 ->
[-1, 1], like dataloaders/sc.py), or `train.synthetic=true` for uniform noise clips on a box without data.  wandb
logging and the periodic sample generation of the reference's loop are not reproduced."""
import os
import sys
import time

import numpy as np
import torch

import diffwave_sashimi_b200 as dwb
from diffwave_sashimi_b200 import distributed as D
from diffwave_sashimi_b200 import experiment as E
from diffwave_sashimi_b200.config import compose
from diffwave_sashimi_b200.training import Trainer

ROOT = os.path.dirname(os.path.abspath(__file__))


class WavFolder(torch.utils.data.Dataset):
    """Every .wav below `root` as a (1, segment_length) float clip in [-1, 1]."""

    def __init__(self, root, segment_length, sampling_rate):
        self.files = sorted(os.path.join(d, f) for d, _, fs in os.walk(root) for f in fs if f.endswith(".wav"))
        if not self.files:
            raise FileNotFoundError(f"no .wav files under {root} (use train.synthetic=true for a data-free run)")
        self.L, self.sr = segment_length, sampling_rate

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        from scipy.io.wavfile import read
        sr, data = read(self.files[i])
        if sr != self.sr:
            raise ValueError(f"{self.files[i]}: {sr} SR doesn't match target {self.sr} SR")
        x = torch.from_numpy(np.asarray(data)).float()
        if np.issubdtype(np.asarray(data).dtype, np.integer):
            x = x / 32768.0
        x = x[: self.L]
        return torch.nn.functional.pad(x, (0, self.L - x.numel())).unsqueeze(0)


class Synthetic(torch.utils.data.Dataset):
    def __init__(self, n, segment_length, seed=0):
        self.n, self.L, self.seed = n, segment_length, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        return torch.rand(1, self.L, generator=g) * 2 - 1


def train(rank, world, diffusion_cfg, model_cfg, dataset_cfg, name=None, ckpt_iter="max", n_iters=1000001, iters_per_ckpt=10000,
          iters_per_logging=100, learning_rate=2e-4, batch_size_per_gpu=4, synthetic=False, seed=0, **_):
    torch.cuda.set_device(rank % torch.cuda.device_count())
    torch.manual_seed(seed + rank)                     # the CPU generator draws diffusion steps and z (train.py:217-218)
    local_path, ckpt_dir = E.local_directory(name, model_cfg, diffusion_cfg, dataset_cfg, "checkpoint")
    dh = dwb.calc_diffusion_hyperparams(**{k: diffusion_cfg[k] for k in ("T", "beta_0", "beta_T", "beta")}, fast=False)
    L = dataset_cfg["segment_length"]
    data = Synthetic(64 * batch_size_per_gpu * world, L, seed) if synthetic else WavFolder(dataset_cfg["data_path"], L, dataset_cfg["sampling_rate"])
    sampler = torch.utils.data.distributed.DistributedSampler(data, world, rank) if world > 1 else None
    loader = torch.utils.data.DataLoader(data, batch_size=batch_size_per_gpu, sampler=sampler, shuffle=sampler is None,
                                         num_workers=0 if synthetic else 4, drop_last=True)
    net = dwb.construct_model(model_cfg).cuda()
    trainer = Trainer(net, batch_size_per_gpu, L, lr=learning_rate)
    it0 = E.find_max_epoch(ckpt_dir) if ckpt_iter == "max" else int(ckpt_iter)
    if it0 >= 0:
        ck = torch.load(os.path.join(ckpt_dir, f"{it0}.pkl"), map_location="cpu")
        with torch.no_grad():
            for k, p in net.named_parameters():      # copy INTO the flat buffer's views (load_state_dict would do the same)
                p.copy_(ck["model_state_dict"][k])
        if "optimizer_state_dict" in ck:
            trainer.load_state_dict(ck["optimizer_state_dict"])
            trainer.lr = float(learning_rate)            # the reference resets the learning rate too (train.py:104-105)
        print(f"Successfully loaded model at iteration {it0}")
    else:
        print("No valid checkpoint model found - training from scratch.")
    if world > 1:
        trainer.broadcast_parameters(0)
    n_iter, t0, clips = it0 + 1, time.time(), 0
    while n_iter < n_iters + 1:
        if sampler is not None:
            sampler.set_epoch(n_iter)
        for audio in loader:
            loss = trainer.loss_backward(audio.cuda(non_blocking=True), dh)
            if world > 1:
                trainer.allreduce_gradients()
            trainer.step()
            clips += audio.shape[0] * world
            if n_iter % iters_per_logging == 0:
                if world > 1:
                    torch.distributed.all_reduce(loss)
                    loss /= world
                if rank == 0:
                    print(f"iteration: {n_iter} \tloss: {float(loss):.6f} \t{clips / (time.time() - t0):.1f} clips/s")
            if n_iter % iters_per_ckpt == 0 and rank == 0:
                torch.save({"model_state_dict": {k: v.detach().cpu().clone() for k, v in net.state_dict().items()},
                            "optimizer_state_dict": trainer.state_dict()}, os.path.join(ckpt_dir, f"{n_iter}.pkl"))
                print(f"model at iteration {n_iter} is saved")
            n_iter += 1
            if n_iter >= n_iters + 1:
                break
    return net, trainer


def main(argv=None):
    cfg = compose(os.path.join(ROOT, "configs"), "config", list(sys.argv[1:] if argv is None else argv))
    rank, world = D.init()
    train(rank, world, cfg.diffusion, cfg.model, cfg.dataset, **cfg.train)


if __name__ == "__main__":
    main()
