"""Per-kernel device times of the SaShiMi hot path at the bench shapes (unet d64, L=16000, batch B):
fftconv and channel mixing of each UNet stage, CUDA events over `iters` back-to-back launches on
rotating buffers (working set > L2).  Used for A/B runs of kernel variants; not a bench value."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffwave_sashimi_b200 as dwb
from oracle.refshim import MODEL_CFGS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = dict(MODEL_CFGS["unet_d64"], n_layers=1)
sd = dwb.init.seeded_state_dict(cfg, seed=0)
net = dwb.construct_model(dict(cfg)); net.load_state_dict(sd); net = net.cuda().eval()
eng = net._engine_get()
ks = eng.s4_kernels()
NB = 3          # rotating buffer sets: 3 x (g, x, out) x B*H*l*4 bytes >> 126 MB L2 at B = 32
def timeit(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for blk, H, l in [(0, 64, 16000), (1, 128, 4000), (2, 256, 1000)]:
    xs = [torch.randn(B, H, l, device="cuda") for _ in range(NB)]
    gs = [torch.randn(B, H, l, device="cuda") for _ in range(NB)]
    st = torch.stack([torch.randn(B, l, device="cuda") * 0.1, torch.rand(B, l, device="cuda") + 0.5], -1).contiguous()
    part = torch.randn(H, device="cuda")
    kf = dwb.ops.fftconv_prepare(ks[blk], torch.randn(H, device="cuda"))
    t_fft = timeit(lambda i: dwb.ops.fftconv(xs[i % NB], kf, st, part, ln_m=0.0, ln_s=1.0))
    t_mix = timeit(lambda i: eng.mix_block(blk, gs[i % NB], xs[i % NB]))
    by = 4.0 * B * H * l
    print(f"stage {blk} H={H} l={l} B={B}: fftconv {t_fft:8.1f} us ({2 * by / t_fft / 1e3:7.1f} GB/s)   mix {t_mix:8.1f} us ({3 * by / t_mix / 1e3:7.1f} GB/s)")
