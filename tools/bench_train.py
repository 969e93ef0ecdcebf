"""Device time of the native WaveNet training step (loss + backward + Adam) on synthetic data:
    python tools/bench_train.py [wnet_h128_d30|wnet_h256_d36] [B] [steps] [mma|simt]
Prints ms per step, clips/s and the fp32 FLOP rate (3x the forward's algorithmic flops)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import diffwave_sashimi_b200 as dwb
from diffwave_sashimi_b200.training import Trainer
from oracle.refshim import MODEL_CFGS          # config table only

name = sys.argv[1] if len(sys.argv) > 1 else "wnet_h128_d30"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gemm = sys.argv[4] if len(sys.argv) > 4 else None
L = 16000
cfg = dict(MODEL_CFGS[name])
net = dwb.construct_model(dict(cfg))
net.load_state_dict(dwb.init.seeded_state_dict(dict(cfg), seed=0))
net = net.cuda().train()
tr = Trainer(net, B, L, gemm=gemm)
dh = dwb.calc_diffusion_hyperparams(200, 1e-4, 0.02)
g = torch.Generator().manual_seed(0)
audio = (torch.rand(B, 1, L, generator=g) * 2 - 1).cuda()
z = torch.randn(B, 1, L, generator=g)
t = torch.randint(200, (B,), generator=g)
losses = []
for _ in range(2):
    losses.append(float(tr.loss_backward(audio, dh, diffusion_steps=t, z=z)))
    tr.step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    loss = tr.loss_backward(audio, dh, diffusion_steps=t, z=z)
    tr.step()
e1.record()
torch.cuda.synchronize()
losses.append(float(loss))
ms = e0.elapsed_time(e1) / steps
C, S, N = cfg["res_channels"], cfg["skip_channels"], cfg["num_res_layers"]
fwd = N * (12 * C * C * L + 2 * C * C * L + 2 * C * S * L) + 2 * S * S * L + 2 * S * L + 2 * C * L
print(json.dumps({"config": name, "gemm": gemm or "default", "B": B, "L": L, "ms_per_step": round(ms, 2), "clips_per_s": round(B / ms * 1e3, 3),
                  "tflops_fp32": round(3 * fwd * B / ms / 1e9, 2), "losses": [round(x, 6) for x in losses], **tr.info()}))
