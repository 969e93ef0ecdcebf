"""Phase timeline of the tcgen05 WaveNet layer (wave_umma.cu) from per-CTA clock64 stamps: python tools/trace_wave.py [C] [layer]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, diffwave_sashimi_b200 as dwb
from diffwave_sashimi_b200._lib import check, lib, ptr, stream_ptr

name = "wnet_h256_d36" if (len(sys.argv) < 2 or sys.argv[1] == "256") else "wnet_h128_d30"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = bench.CONFIGS[name]["cfg"]
net = bench.build_net(name, torch.device("cuda"))
eng = net._engine_get()
B, L, C, S = 8, 16000, cfg["res_channels"], cfg["skip_channels"]
h = torch.randn(B, C, L, device="cuda"); part = torch.randn(C, device="cuda")
ho = torch.empty_like(h); skip = torch.zeros(B, S, L, device="cuda")
nt = B * ((L + 127) // 128)
tr = torch.zeros(nt, 16, dtype=torch.int64, device="cuda")
for _ in range(3):
    check(lib().dwb_debug_wave_trace(eng._plan, layer, ptr(h), ptr(part), ptr(ho), ptr(skip), B, L, ptr(tr), stream_ptr(h.device)))
torch.cuda.synchronize()
t = tr.cpu().numpy().astype(np.float64)
names = {1: "setup done (barriers, TMEM alloc)", 15: "loader: first slab stored", 13: "MMA: first slab seen", 14: "MMA: first weight stage seen", 8: "MMA: first K chunk issued", 2: "loaders done", 9: "MMA: phase 1 issued",
         3: "E1 start (acc1 ready)", 4: "E1 done", 5: "E2 chunk 0 ready", 10: "MMA: phase-2 chunk 0 issued", 11: "MMA: all issued",
         6: "E2 last chunk ready", 7: "E2 done", 12: "exit"}
print(f"{name} layer {layer} (dilation {2 ** (layer % cfg['dilation_cycle'])}), {nt} tiles; cycles since CTA start (median over CTAs)")
for k in (1, 15, 13, 14, 8, 2, 9, 3, 4, 5, 10, 11, 6, 7, 12):
    d = t[:, k] - t[:, 0]
    print(f"  {names[k]:38s} {np.median(d):9.0f}   (p10 {np.percentile(d, 10):8.0f}  p90 {np.percentile(d, 90):8.0f})")
