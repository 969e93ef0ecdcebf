"""Multi-GPU generation: independent clips sharded over one process per GPU, no collective inside
the T-step loop, one all_gather at sample collection.

The reference launches one independent process per GPU with no communication at all
(generate.py:217-227) and every process draws from its own unseeded generator.  Here every clip of
the GLOBAL batch owns a generator seeded from (seed, global clip index) and consumed in the
reference's order, so a rank draws only its own clips and 1-GPU and N-GPU runs produce identical clips.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Join the process group described by torchrun's environment (RANK/WORLD_SIZE/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, **kw)
    return int(os.environ.get("RANK", "0")), world


def shard_range(n_clips, rank, world):
    """Contiguous shard [lo, hi) of rank; the first n_clips % world ranks take one extra clip."""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def draw_noise_sharded(global_size, T, seed, rank, world):
    """(x_T, noise) for this rank's clips [lo, hi) of a global batch `global_size` = (B, 1, L), materialised.
    Clip c draws from its own generator seeded with clip_seed(seed, c) in the reference's order (x_T, then one
    draw per step, generate.py:47,54), so a rank draws only its own rows and 1-GPU and N-GPU runs give
    identical clips.  `generate_sharded` streams the same draws instead of materialising them."""
    from .sampler import PerClipNoise, clip_seed
    B = global_size[0]
    lo, hi = shard_range(B, rank, world)
    src = PerClipNoise([clip_seed(seed, c) for c in range(lo, hi)])
    x_T = torch.empty((hi - lo,) + tuple(global_size[1:]))
    noise = torch.empty((max(T - 1, 0), hi - lo) + tuple(global_size[1:]))
    if hi > lo:
        src.fill(x_T)
        for i in range(T - 1):
            src.fill(noise[i])
    return x_T, noise


def gather_samples(local, n_clips, rank, world):
    """all_gather of the per-rank (b_r, 1, L) results into the global (B, 1, L) batch — the only
    collective of the generation path.  Ragged shards are padded to the largest shard."""
    if world == 1:
        return local
    sizes = [shard_range(n_clips, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)])


@torch.no_grad()
def generate_sharded(net, n_clips, L, diffusion_hyperparams, seed, condition=None, rank=0, world=1):
    """Global batch of n_clips through `sampling`, sharded over ranks; returns the full (n_clips, 1, L) batch on
    every rank.  Per-clip noise streams (sampler.PerClipNoise): the result does not depend on `world`."""
    from .sampler import PerClipNoise, clip_seed, sampling
    lo, hi = shard_range(n_clips, rank, world)
    eng = net._engine_get()
    if hi > lo:
        src = PerClipNoise([clip_seed(seed, c) for c in range(lo, hi)])
        local = sampling(net, (hi - lo, 1, L), diffusion_hyperparams, condition, verbose=False, noise=src)
    else:
        local = torch.empty((0, 1, L), device=eng.device)
    return gather_samples(local, n_clips, rank, world)


# ---- data-parallel training (distributed_util.py:97-149) ---------------------------------------------------------
def broadcast_flat(params, src=0):
    """Rank `src`'s flat parameter buffer to every rank: ONE broadcast where the reference broadcasts each
    state_dict tensor separately (distributed_util.py:107-110)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(params, src)


def allreduce_flat(grads, average=True):
    """Sum (or mean) of the flat gradient buffer over ranks, in place, in ONE all-reduce; returns the world size.
    The reference flattens every gradient into a scratch tensor, all-reduces, divides and copies back per step
    (distributed_util.py:119-138); the trainer's gradients already are one contiguous buffer, and with
    average=False the 1/world_size is left to the fused Adam kernel (Trainer.step)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(grads)
        if average:
            grads /= world
    return world
