"""GPU debug: per-CTA phase timeline of the tcgen05 mixing kernel (clock64 deltas, cycles)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffwave_sashimi_b200 as dwb
from diffwave_sashimi_b200._lib import lib, ptr, stream_ptr, check
from oracle.refshim import MODEL_CFGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = dict(MODEL_CFGS["unet_d64"], n_layers=1)
sd = dwb.init.seeded_state_dict(cfg, seed=0)
net = dwb.construct_model(dict(cfg)); net.load_state_dict(sd); net = net.cuda().eval()
eng = net._engine_get()
L = lib()
L.dwb_debug_mix_trace.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
L.dwb_debug_mix_trace.restype = ctypes.c_int
pers = os.environ.get("DWB_UMMA") != "tile"
names = ["top", "g_arrive", "x_tmem", "acc1", "E1", "z_arrive", "acc2", "E2", "acc3", "E3end"] if pers else ["start", "setup", "g_arrive", "x_tmem", "acc1", "E1", "z_arrive", "acc2", "E2", "acc3", "E3", "endsync", "mma_g", "mma_z", "mma_end"]
for blk, H, l in [(0, 64, 16000), (1, 128, 4000)]:
    g = torch.randn(B, H, l, device="cuda"); x = torch.randn(B, H, l, device="cuda")
    out = torch.empty_like(x); st = torch.empty(B, l, 2, device="cuda")
    ntile = (l + 127) // 128
    tr = torch.zeros(B * ntile, 16, dtype=torch.int64, device="cuda")
    for it in range(3):
        check(L.dwb_debug_mix_trace(eng._plan, blk, ptr(g), ptr(x), ptr(out), ptr(st), B, ptr(tr), stream_ptr()))
    torch.cuda.synchronize()
    t = tr.cpu().double()
    t = t[t[:, 1] > 0]
    rel = t[:, :15] - t[:, :1]
    print(f"H={H}: {B*ntile} CTAs; median cycles since CTA start:")
    med = rel.median(0).values
    for i, n in enumerate(names):
        print(f"   {n:10s} {med[i]:9.0f}")
    if pers:
        continue
    print("   total per CTA: median %.0f  p10 %.0f  p90 %.0f" % (rel[:, 11].median(), rel[:, 11].kthvalue(max(1, int(0.1 * len(rel)))).values, rel[:, 11].kthvalue(int(0.9 * len(rel))).values))
