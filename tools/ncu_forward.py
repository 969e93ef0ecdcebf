"""Short workload for ncu: a few eager forwards of the bench model (unet d64, L=16000) at batch B.
Profiling is limited to the region between cudaProfilerStart/Stop (use --profile-from-start off).
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_forward.py --batch 16
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import diffwave_sashimi_b200 as dwb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--model", default="unet_d64")
ap.add_argument("--forwards", type=int, default=1)
args = ap.parse_args()

cfg = dict(bench.CFG)
if args.model == "wnet_h256_d36":
    cfg = dict(_name_="wavenet", unconditional=True, in_channels=1, out_channels=1, res_channels=256, skip_channels=256,
               num_res_layers=36, dilation_cycle=12, diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
               diffusion_step_embed_dim_out=512)
elif args.model == "wnet_h128_d30":
    cfg = dict(_name_="wavenet", unconditional=True, in_channels=1, out_channels=1, res_channels=128, skip_channels=256,
               num_res_layers=30, dilation_cycle=10, diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
               diffusion_step_embed_dim_out=512)
sd = dwb.init.seeded_state_dict(cfg, seed=0)
net = dwb.construct_model(dict(cfg))
net.load_state_dict(sd)
net = net.cuda().eval()
g = torch.Generator().manual_seed(1)
x = torch.randn(args.batch, 1, 16000, generator=g).cuda()
t = torch.full((args.batch, 1), 100.0).cuda()
with torch.no_grad():
    for _ in range(2):
        net((x, t))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.forwards):
        net((x, t))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
