"""Multi-GPU generation: independent clips sharded over one process per GPU, no collective inside
the T-step loop, one all_gather at sample collection.

The reference launches one independent process per GPU with no communication at all
(generate.py:217-227) and every process draws from its own unseeded generator.  Here the noise of
the GLOBAL batch is defined once (reference draw order on one seeded CPU generator) and each rank
takes its contiguous slice, so 1-GPU and N-GPU runs produce identical clips.
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


def draw_noise_sharded(global_size, T, seed, rank, world, chunk=None):
    """(x_T, noise) for this rank's clips of a global batch `global_size` = (B, 1, L).
    The stream is the reference's (x_T, then one (B,1,L) draw per step, generate.py:47,54) on a CPU
    generator seeded with `seed`; every rank draws the same stream and keeps rows [lo, hi)."""
    B = global_size[0]
    lo, hi = shard_range(B, rank, world)
    g = torch.Generator().manual_seed(seed)
    x_T = torch.normal(0, 1, size=global_size, generator=g)[lo:hi].clone()
    noise = torch.empty((max(T - 1, 0), hi - lo) + tuple(global_size[1:]))
    for i in range(T - 1):
        noise[i] = torch.normal(0, 1, size=global_size, generator=g)[lo:hi]
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
    """Global batch of n_clips through `sampling` semantics, sharded over ranks; returns the full
    (n_clips, 1, L) batch on every rank."""
    from .sampler import step_coefficients
    eng = net._engine_get()
    T = diffusion_hyperparams["T"]
    x_T, noise = draw_noise_sharded((n_clips, 1, L), T, seed, rank, world)
    local = eng.sample(x_T.to(eng.device), noise.to(eng.device), step_coefficients(diffusion_hyperparams), condition)
    return gather_samples(local, n_clips, rank, world)
