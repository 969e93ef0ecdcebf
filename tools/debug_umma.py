"""GPU debug: tcgen05 channel mixing (mix_umma.cu) vs the exact fp32 SIMT kernel, block by block."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffwave_sashimi_b200 as dwb
from oracle.refshim import MODEL_CFGS

def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = dict(MODEL_CFGS["unet_d64"], L=L, n_layers=1)
sd = dwb.init.seeded_state_dict(cfg, seed=0)
net = dwb.construct_model(dict(cfg)); net.load_state_dict(sd); net = net.cuda().eval()
eng = net._engine_get()
g = torch.Generator().manual_seed(3)
blocks = [(0, 64, L), (1, 128, L // 4), (2, 256, L // 16)]
for blk, H, l in blocks:
    gg = torch.randn(B, H, l, generator=g).cuda()
    x = torch.randn(B, H, l, generator=g).cuda() * 1.5 + 0.3
    sk = torch.randn(B, H, l, generator=g).cuda()
    for skip in (None, sk):
        o1, s1 = eng.mix_block(blk, gg, x, skip, exact=True)
        torch.cuda.synchronize()
        o2, s2 = eng.mix_block(blk, gg, x, skip, exact=False)
        torch.cuda.synchronize()
        print(f"block {blk} H={H} l={l} skip={skip is not None}: out rel {rel(o2, o1):.3e} stats rel {rel(s2, s1):.3e} "
              f"max|d| {(o2 - o1).abs().max().item():.3e} finite={bool(torch.isfinite(o2).all())}")
        if rel(o2, o1) > 1e-3:
            d = (o2 - o1).abs()
            print("  err by row-in-tile (first tile):", d[0, :, :128].amax(0)[:16].tolist())
            print("  err by channel:", d[0].amax(1)[:16].tolist())
