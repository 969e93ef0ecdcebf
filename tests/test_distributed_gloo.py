"""CPU, world_size 2, gloo: sharding, rank-independent noise and the single gather of the N>1
generation path (the per-rank engine needs a GPU; a stand-in sampler that consumes the same
inputs stands in for it here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sampler(x_T, noise):
    # any deterministic per-clip function of exactly the tensors the real sampler consumes
    w = torch.linspace(0.5, 1.5, noise.shape[0]).view(-1, 1, 1, 1) if noise.shape[0] else noise.new_zeros(0, 1, 1, 1)
    return torch.tanh(x_T) + (noise * w).sum(0) * 0.01


def _worker(rank, world, port, n_clips, L, T, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from diffwave_sashimi_b200 import distributed as D
    r, w = D.init("gloo")
    assert (r, w) == (rank, world)
    x_T, noise = D.draw_noise_sharded((n_clips, 1, L), T, seed, rank, world)
    lo, hi = D.shard_range(n_clips, rank, world)
    assert x_T.shape == (hi - lo, 1, L) and noise.shape == (T - 1, hi - lo, 1, L)
    full = D.gather_samples(_fake_sampler(x_T, noise), n_clips, rank, world)
    # device-time style reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((full, t.item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [4, 5])
def test_two_ranks_match_one(n_clips):
    from diffwave_sashimi_b200 import distributed as D
    from oracle import diffwave_oracle as O
    L, T, seed, world = 64, 6, 77, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, L, T, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run of the same global batch; every clip's stream is the reference's draw order for a
    # batch of ONE clip under that clip's seed (oracle restatement of generate.py:47,54)
    from diffwave_sashimi_b200.sampler import clip_seed
    x_T, noise = D.draw_noise_sharded((n_clips, 1, L), T, seed, 0, 1)
    for c in range(n_clips):
        ox, on = O.draw_noise(clip_seed(seed, c), (1, 1, L), T)
        assert torch.equal(x_T[c:c + 1], ox) and torch.equal(noise[:, c:c + 1], on)
    assert torch.equal(full, _fake_sampler(x_T, noise))
    assert tmax == 2.0


def test_shard_ranges_cover_batch():
    from diffwave_sashimi_b200 import distributed as D
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def _train_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from diffwave_sashimi_b200 import distributed as D
    D.init("gloo")
    g = torch.Generator().manual_seed(100 + rank)
    params = torch.randn(n, generator=g)
    grads = torch.randn(n, generator=g)
    D.broadcast_flat(params, 0)
    summed = grads.clone()
    w = D.allreduce_flat(summed, average=False)
    mean = grads.clone()
    D.allreduce_flat(mean, average=True)
    if rank == 1:
        q.put((params, summed, mean, w))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    """The training path's two collectives on the flat buffers (the reference's apply_gradient_allreduce,
    distributed_util.py:97-149): parameters follow rank 0, gradients are averaged."""
    n, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    params, summed, mean, w = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    p0, g0 = torch.randn(n, generator=gens[0]), torch.randn(n, generator=gens[0])
    _, g1 = torch.randn(n, generator=gens[1]), torch.randn(n, generator=gens[1])
    assert w == 2 and torch.equal(params, p0)
    assert torch.allclose(summed, g0 + g1) and torch.allclose(mean, (g0 + g1) / 2)
