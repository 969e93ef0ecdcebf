"""Experiment: do two half-batch sampler graphs on two streams overlap better than one full-batch graph?
    python tools/bench_dual.py [--batch 32] [--T 20] [--ways 2]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import diffwave_sashimi_b200 as dwb

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--T", type=int, default=20)
ap.add_argument("--ways", type=int, default=2)
ap.add_argument("--delay", type=int, default=0, help="stagger: stream i starts after i*delay forwards of spin")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, T, L, W = args.batch, args.T, 16000, args.ways
sd = dwb.init.seeded_state_dict(bench.CFG, seed=0)
nets = []
for _ in range(W + 1):
    net = dwb.construct_model(dict(bench.CFG)); net.load_state_dict(sd); nets.append(net.cuda().eval())
dh = dwb.calc_diffusion_hyperparams(T, 1e-4, 0.02, fast=True)
coef = dwb.step_coefficients(dh)
torch.manual_seed(0)
x_T = torch.randn(B, 1, L, device=dev); noise = torch.randn(T - 1, B, 1, L, device=dev)
out = torch.empty_like(x_T)
full = nets[0]._engine_get()
def t_ms(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
single = t_ms(lambda: full.sample(x_T, noise, coef, out=out))
streams = [torch.cuda.Stream() for _ in range(W)]
Bs = B // W
xs = [x_T[i * Bs:(i + 1) * Bs].contiguous() for i in range(W)]
ns = [noise[:, i * Bs:(i + 1) * Bs].contiguous() for i in range(W)]
outs = [torch.empty_like(x) for x in xs]
engs = [nets[i + 1]._engine_get() for i in range(W)]
def dual():
    cur = torch.cuda.current_stream()
    for i, s in enumerate(streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            if args.delay and i:
                torch.cuda._sleep(int(args.delay * i))
            engs[i].sample(xs[i], ns[i], coef, out=outs[i])
    for s in streams:
        cur.wait_stream(s)
multi = t_ms(dual)
ref = torch.cat(outs, 0)
err = ((ref - out).norm() / out.norm()).item()
print(json.dumps({"B": B, "T": T, "ways": W, "delay": args.delay, "single_ms_per_step": single / T, "multi_ms_per_step": multi / T,
                  "speedup": single / multi, "rel_diff": err, "fft": os.environ.get("DWB_FFT", "default")}))
