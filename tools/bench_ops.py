"""Per-op timing + numerical check of the two hot kernels at the bench shapes (unet d64, B clips).

    python tools/bench_ops.py [--batch 32] [--iters 20] [--ops fft,mix]

fftconv is timed through the single-op C ABI (dwb_fftconv) and checked against a float64 torch.fft
evaluation of the same formula (tool-side checker only); the channel mixing is timed through
dwb_plan_mix_block and checked against the exact-fp32 SIMT kernel.  Activation buffers are rotated
so every launch streams from HBM (total > L2).  One JSON line per (op, stage).
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import diffwave_sashimi_b200 as dwb  # noqa: E402
from diffwave_sashimi_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--ops", default="fft,mix")
ap.add_argument("--tag", default=os.environ.get("DWB_TAG", ""))
args = ap.parse_args()
dev = torch.device("cuda", 0)
B = args.batch
STAGES = [(64, 16000), (128, 4000), (256, 1000)]


def timeit(fn, nbuf):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters * 1e3      # us


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


if "fft" in args.ops:
    for (H, l) in STAGES:
        g = torch.Generator(device=dev).manual_seed(H)
        k = torch.randn(2, H, l, generator=g, device=dev) * torch.exp(-torch.arange(l, device=dev) / (l / 8.0))
        D = torch.randn(H, generator=g, device=dev)
        kf = ops.fftconv_prepare(k, D)
        nbuf = max(2, int(math.ceil(400e6 / (B * H * l * 4))))
        xs = [torch.randn(B, H, l, generator=g, device=dev) for _ in range(nbuf)]
        stats = torch.stack([xs[0].mean(1), 1.0 / xs[0].std(1, unbiased=False)], -1).contiguous()      # (B,l,2)
        pt = torch.randn(B, H, generator=g, device=dev)
        # float64 check on a slice of the batch
        nb = min(B, 2)
        y = (1.3 * stats[:nb, :, 1].unsqueeze(1).double()) * (xs[0][:nb].double() - stats[:nb, :, 0].unsqueeze(1).double() + 0.2) \
            + pt[:nb].double().unsqueeze(-1)
        n = 2 * l
        kk = torch.nn.functional.pad(k[0].double(), (0, l)) + torch.nn.functional.pad(k[1].double().flip(-1), (l, 0))
        c = torch.fft.irfft(torch.fft.rfft(y, n=n) * torch.fft.rfft(kk, n=n), n=n)[..., :l] + D.double()[None, :, None] * y
        ref = torch.nn.functional.gelu(c)
        got = ops.fftconv(xs[0][:nb].contiguous(), kf, stats=stats[:nb].contiguous(), part_t=pt[:nb].contiguous(), ln_m=0.2, ln_s=1.3)
        err = rel(got, ref)
        us = timeit(lambda i: ops.fftconv(xs[i], kf, stats=stats, part_t=pt, ln_m=0.2, ln_s=1.3), nbuf)
        us_ns = timeit(lambda i: ops.fftconv(xs[i], kf, stats=None, part_t=pt), nbuf)      # no LayerNorm statistics to read
        nbytes = 2 * 4.0 * B * H * l
        print(json.dumps({"tag": args.tag, "op": "fftconv", "H": H, "l": l, "B": B, "us": round(us, 1), "us_without_stats": round(us_ns, 1),
                          "GBs": round(nbytes / us / 1e3, 1), "rel_l2_vs_f64": err}))

if "mix" in args.ops:
    sd = dwb.init.seeded_state_dict(bench.CFG, seed=0)
    net = dwb.construct_model(dict(bench.CFG))
    net.load_state_dict(sd)
    net = net.cuda().eval()
    eng = net._engine_get()
    blocks = {0: 0, 1: 6, 2: 12}       # first block of each stage (d_layers 0.., 7.., c_layers)
    for s, (H, l) in enumerate(STAGES):
        g = torch.Generator(device=dev).manual_seed(s)
        nbuf = max(2, int(math.ceil(300e6 / (B * H * l * 4))))
        gs = [torch.randn(B, H, l, generator=g, device=dev) for _ in range(nbuf)]
        xs = [torch.randn(B, H, l, generator=g, device=dev) for _ in range(nbuf)]
        o_fast, st_fast = eng.mix_block(blocks[s], gs[0][:2], xs[0][:2])
        o_ex, st_ex = eng.mix_block(blocks[s], gs[0][:2], xs[0][:2], exact=True)
        err, serr = rel(o_fast, o_ex), rel(st_fast, st_ex)
        us = timeit(lambda i: eng.mix_block(blocks[s], gs[i], xs[i]), nbuf)
        nbytes = 3 * 4.0 * B * H * l
        print(json.dumps({"tag": args.tag, "op": "mix", "H": H, "l": l, "B": B, "us": round(us, 1),
                          "GBs": round(nbytes / us / 1e3, 1), "rel_l2_vs_simt": err, "stats_rel": serr}))
