"""Per-CTA phase timeline of the fused tcgen05 mixing kernels (clock64 deltas, cycles): python tools/trace_umma.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, diffwave_sashimi_b200 as dwb
from diffwave_sashimi_b200._lib import check, lib, ptr, stream_ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = dict(bench.CFG, n_layers=1)
sd = dwb.init.seeded_state_dict(cfg, seed=0)
net = dwb.construct_model(dict(cfg)); net.load_state_dict(sd); net = net.cuda().eval()
eng = net._engine_get()
pers_env = os.environ.get("DWB_UMMA")
for blk, H, l in [(0, 64, 16000), (1, 128, 4000), (2, 256, 1000)]:
    ntile_all = B * ((l + 127) // 128)
    pers = H != 256 and ((pers_env == "pers") or (pers_env is None and (H == 128 or ntile_all >= 6000)))
    names = (["top", "acc1", "E1", "z_arrive+x_next", "acc2", "E2", "acc3", "g_next", "E3end"] if pers else
             ["start", "setup", "g_arrive", "x_tmem", "acc1", "E1", "z_arrive", "acc2", "skip_ld", "acc3", "E3", "endsync", "mma_g", "mma_z", "mma_end"])
    g = torch.randn(B, H, l, device="cuda"); x = torch.randn(B, H, l, device="cuda")
    out = torch.empty_like(x); st = torch.empty(B, l, 2, device="cuda")
    ntile = (l + 127) // 128
    tr = torch.zeros(B * ntile, 16, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        if it == 2: e0.record()
        check(lib().dwb_debug_mix_trace(eng._plan, blk, ptr(g), ptr(x), ptr(out), ptr(st), B, ptr(tr), stream_ptr(g.device)))
    e1.record()
    torch.cuda.synchronize()
    t = tr.cpu().double()
    t = t[t[:, 1] > 0]
    rel = t[:, :15] - t[:, :1]
    print(f"H={H} l={l} B={B}: {B*ntile} tiles, {'persistent' if pers else 'per-tile'} kernel, {e0.elapsed_time(e1)*1e3:.0f} us; median cycles since CTA/tile start:")
    if len(rel) == 0:
        continue
    med = rel.median(0).values
    for i, n in enumerate(names):
        print(f"   {n:10s} {med[i]:9.0f}")
    if pers:
        raw = t[:, 10:14].median(0).values
        print(f"   MMA issuer of group 0, second tile: waits for g/z operands {raw[0]:.0f}, weights {raw[1]:.0f}, hidden chunks {raw[2]:.0f} of {raw[3]:.0f} cycles (issue to issue)")
