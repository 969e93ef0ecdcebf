#!/usr/bin/env python
"""bench.py — audio clips/sec of the DiffWave reverse-sampling loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--config NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full T-step reverse loop (generate.py:23-55) over one batch of B synthetic clips
per GPU: x_T and the T-1 noise draws go in, x_0 comes out.  The default workload at every N is
BASELINE.json configs[1]: SC09 unconditional SaShiMi unet d64 n6 pool=[4,4] expand=2 ff=2, T=200,
L=16000, random-init weights (seeded HiPPO-LegS initialiser, final conv made non-zero).
`--config` selects one of the other BASELINE.json configs (unet_d128, unet_d32_cond, wnet_h256_d36,
wnet_h128_d30) for the same line; the default line also carries a short measurement of each under
`other_configs` (N=1 only).

Printed JSON (one line, rank 0):
  value      clips/s over all GPUs, inputs resident in HBM, one-step CUDA graph replayed T times per step
  e2e        same metric through the public API `sampling(net, size, dh)` exactly as generate.py calls it:
             x_T and every step's noise are drawn on the CPU generator (reference order) INSIDE the timed region,
             staged through pinned memory, copied host->device chunk by chunk under the running steps, and x_0 is
             read back to the host.  `e2e.predrawn` = the same with the draws made before the timer starts.
  roofline   the dominant kernel (largest share of a forward, timed live with CUDA events around every launch
             on the launching stream: dwb_plan_profile) against the measured peak that bounds it: HBM bytes for the
             SaShiMi kernels (2*4*H*l*B per S4-convolution launch), tensor FLOPs for the WaveNet layer;
             `traffic` = dram read+write bytes per launch from the committed ncu capture (profiles/ncu_traffic.json);
             `whole_loop` = dwb_plan_work bytes (flops) x T x B per step / step time; `kernels` lists every category
  check      finiteness of the timed batch and clip 0 of it against a B=1 run of the same clip
  cpu_baseline  the oracle port of the reference's CPU path (fp32, S4 kernels regenerated every
             step exactly as the reference does), timed on this host for a bounded sample
`--impl reference` runs only that CPU arm (rank 0), same metric/config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_EMB = dict(diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512, diffusion_step_embed_dim_out=512)


def _sashimi(d, cond=False):
    c = dict(_name_="sashimi", unconditional=not cond, in_channels=1, out_channels=1, unet=True, d_model=d, n_layers=6,
             pool=[4, 4], expand=2, ff=2, L=16000, **_EMB)
    if cond:
        c["mel_upsample"] = [16, 16]
    return c


def _wavenet(C, S, N, cycle):
    return dict(_name_="wavenet", unconditional=True, in_channels=1, out_channels=1, res_channels=C, skip_channels=S,
                num_res_layers=N, dilation_cycle=cycle, **_EMB)


# BASELINE.json configs (configs/model/*.yaml + configs/experiment/*.yaml of the reference)
CONFIGS = {
    "unet_d64": dict(cfg=_sashimi(64), T=200, beta_T=0.02, batch=64, mel=None, bound="hbm",
                     workload="SC09 unconditional SaShiMi unet d64 n6 pool=[4,4] expand=2 ff=2 T=200 L=16000"),
    "unet_d128": dict(cfg=_sashimi(128), T=200, beta_T=0.02, batch=8, mel=None, bound="hbm",
                      workload="SC09 unconditional SaShiMi unet d128 n6 T=200 L=16000 (global batch 64 on 8 GPUs = 8 per GPU)"),
    "unet_d32_cond": dict(cfg=_sashimi(32, cond=True), T=50, beta_T=0.05, batch=16, mel=(1, 80, 63), bound="hbm",
                          workload="LJSpeech mel-conditioned vocoder SaShiMi d32 n6 T=50 L=16000 hop=256"),
    "wnet_h256_d36": dict(cfg=_wavenet(256, 256, 36, 12), T=200, beta_T=0.02, batch=8, mel=None, bound="tensor",
                          workload="SC09 unconditional WaveNet h256/d36 T=200 L=16000"),
    "wnet_h128_d30": dict(cfg=_wavenet(128, 256, 30, 10), T=200, beta_T=0.02, batch=8, mel=None, bound="tensor",
                          workload="SC09 unconditional WaveNet h128/d30 T=200 L=16000"),
}
DEFAULT = "unet_d64"
CFG = CONFIGS[DEFAULT]["cfg"]            # tools/ import these
WORKLOAD = CONFIGS[DEFAULT]["workload"]
BETA_0, L = 1e-4, 16000
def kernel_label(kname, cfg, B):
    """Which kernel a profile category runs for this config (the dispatch rules of csrc/api.cu, mix_umma.cu, fftconv.cu)."""
    if kname == "wave_block":
        C = cfg["res_channels"]
        return ("wave_block_umma_kernel<%d,%d> (tcgen05 residual layer)" % (C, cfg["skip_channels"])) if C in (128, 256) \
            else "wave_block_mma_kernel (mma.sync)"
    if kname[:-1] in ("fftconv_s", "mix_s"):
        s_ = int(kname[-1])
        H, l = cfg["d_model"] * cfg["expand"] ** s_, L // (4 ** s_)
        if kname.startswith("fft"):
            lg = max(4, (l - 1).bit_length())
            return f"fftconv3_kernel<{lg}> (H={H}, l={l})" if lg == 14 and l % 4 == 0 else f"fftconv_kernel<{lg}> (H={H}, l={l})"
        k = {64: "sashimi_mix_umma_pers_kernel<64,2>", 128: "sashimi_mix_umma_pers_kernel<128,2>",
             256: "sashimi_mix_umma256_kernel"}.get(H)
        if k is None:
            k = ("pool_umma_kernel<GLU|GELU|RES> + channel_stats_kernel" if H == 512 else "mix_gemm_umma_kernel x3 + channel_stats_kernel x2") \
                if H % 128 == 0 else f"sashimi_mix_mma_kernel<{H}> (mma.sync)"
        return f"{k} (H={H}, l={l})"
    return kname


def metric_name(T):
    return f"audio clips/sec (16k-sample, T={T})"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's own CPU path
# --------------------------------------------------------------------------------------------
def cpu_arm(name, steps, warmup, budget_s=25.0):
    """Time `steps` single diffusion steps (B=1) of the reference algorithm on the host cores:
    fp32, all threads, S4 kernels regenerated inside every step like models/s4.py:1388 does.
    clips/s = 1 / (T * mean step seconds).  Also times the same step with hoisted kernels."""
    import torch
    from oracle import diffwave_oracle as O
    import diffwave_sashimi_b200 as dwb
    spec = CONFIGS[name]
    cfg, T = spec["cfg"], spec["T"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = dwb.init.seeded_state_dict(cfg, seed=0)
    sash = cfg["_name_"] == "sashimi"
    if sash:
        lay = O.sashimi_layout(cfg)
        for sec in "dcu":                      # the reference does this rewrite on its first forward
            for (p, kind, H, l, _) in lay[sec]:
                if kind == "block":
                    sd[p + "layer.kernel.kernel.C"] = O.s4_setup_C(sd, p + "layer.", l).float()
                    sd[p + "layer.kernel.kernel.L"] = torch.tensor(l)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 1, L, generator=g)
    mel = torch.randn(*spec["mel"], generator=g) if spec["mel"] else None
    t = torch.full((1, 1), float(T // 2))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.forward(cfg, sd, x, t, mel=mel, dtype=torch.float32)  # kernels=None -> regenerated, as shipped
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if i >= warmup and sum(times) > budget_s:
                break
        hoisted = None
        if sash:
            ks = O.sashimi_kernels(cfg, sd, dtype=torch.float32)
            t0 = time.perf_counter()
            O.forward(cfg, sd, x, t, mel=mel, dtype=torch.float32, kernels=ks)
            hoisted = time.perf_counter() - t0
    mean = sum(times) / len(times)
    out = {"value": 1.0 / (T * mean), "unit": "clips/s", "cores": cores, "kind": "port",
           "sample": f"{len(times)} diffusion steps of {T} at B=1 (mean {mean:.2f} s/step), extrapolated x{T}"
                     + ("; S4 kernels regenerated every step as the reference does" if sash else ""),
           "s_per_step": mean}
    if hoisted is not None:
        out["s_per_step_kernels_hoisted"] = hoisted
        out["value_kernels_hoisted"] = 1.0 / (T * hoisted)
    return out, len(times), mean


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = CONFIGS[args.config]
    cb, n, mean = cpu_arm(args.config, args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": metric_name(spec["T"]), "value": cb["value"], "unit": "clips/s", "n_gpus": args.gpus,
            "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": mean * 1e3 * spec["T"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": spec["workload"], "batch": 1, "note": "reference CPU path (oracle port), bounded sample"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def build_net(name, dev):
    import diffwave_sashimi_b200 as dwb
    spec = CONFIGS[name]
    sd = dwb.init.seeded_state_dict(spec["cfg"], seed=0)      # identical weights on every rank
    net = dwb.construct_model(dict(spec["cfg"]))
    net.load_state_dict(sd)
    return net.to(dev).eval()


def measure_training(name, B, dev, steps=2, warmup=1):
    """Device time of the native training step (loss + backward + Adam, csrc/train_wavenet.cu) on
    config `name` at B clips of L samples; fp32 flops = 3x the forward's algorithmic flops."""
    import torch
    import diffwave_sashimi_b200 as dwb
    from diffwave_sashimi_b200.training import Trainer
    spec = CONFIGS[name]
    net = build_net(name, dev).train()
    tr = Trainer(net, B, L)
    dh = dwb.calc_diffusion_hyperparams(spec["T"], BETA_0, spec["beta_T"])
    g = torch.Generator().manual_seed(99)
    audio = (torch.rand(B, 1, L, generator=g) * 2 - 1).to(dev)
    z, t = torch.randn(B, 1, L, generator=g).to(dev), torch.randint(spec["T"], (B,), generator=g)
    for _ in range(warmup):
        tr.loss_backward(audio, dh, diffusion_steps=t, z=z)
        tr.step()
    n0 = tr.info()["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        loss = tr.loss_backward(audio, dh, diffusion_steps=t, z=z)
        tr.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    c = spec["cfg"]
    C, S, N = c["res_channels"], c["skip_channels"], c["num_res_layers"]
    fwd = N * (12 * C * C * L + 2 * C * C * L + 2 * C * S * L) + 2 * S * S * L + 2 * S * L + 2 * C * L
    out = {"workload": f"one optimizer step (loss + backward + Adam) of {spec['workload'].split(' T=')[0]}, L={L}", "batch": B,
           "ms_per_step": round(ms, 2), "value": round(B / ms * 1e3, 3), "unit": "clips/s per optimizer step", "steps": steps,
           "warmup": warmup, "dtype": "f32", "tflops_fp32": round(3 * fwd * B / ms / 1e9, 2), "loss": round(float(loss), 6),
           "finite": bool(torch.isfinite(loss)), "gpu_launches": (tr.info()["launches"] - n0) // steps + 1,
           "note": "exact-fp32 software-pipelined SIMT GEMM tiles (tensor-core option measured slower); not part of the headline metric"}
    _, tf, src = peaks()
    out["roofline"] = {"bound": "tensor", "achieved": out["tflops_fp32"], "peak": tf, "unit": "TFLOP/s",
                       "frac": round(out["tflops_fp32"] / tf, 4), "peak_source": src,
                       "note": "fp32-equivalent flops against the single-pass bf16 peak the WaveNet rows are judged by; this path runs on the FP32 pipe"}
    tr.close()
    return out


def measure(name, B, K, W, dev, world=1, rank=0, full=True):
    """Resident + end-to-end timing of config `name` at B clips per GPU; full=False: resident timing only."""
    import torch
    import torch.distributed as dist
    import diffwave_sashimi_b200 as dwb
    spec = CONFIGS[name]
    T = spec["T"]
    net = build_net(name, dev)
    eng = net._engine_get()
    dh = dwb.calc_diffusion_hyperparams(T, BETA_0, spec["beta_T"], fast=True)
    coef = dwb.step_coefficients(dh)
    g = torch.Generator().manual_seed(4321)
    mel = torch.randn(*spec["mel"], generator=g).to(dev) if spec["mel"] else None
    torch.manual_seed(1234 + rank)
    x_T_h, noise_h = dwb.draw_noise((B, 1, L), T, pin=True)
    x_T, noise = x_T_h.to(dev), noise_h.to(dev)
    out = torch.empty_like(x_T)
    gathered = [torch.empty_like(out) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def step_resident():
        eng.sample(x_T, noise, coef, mel_spec=mel, out=out)
        if world > 1:                       # the single collective: sample collection
            dist.all_gather(gathered, out)

    # ---- resident-input timing: W warm-up + exactly K timed steps --------------------------
    for _ in range(W):
        step_resident()
    barrier()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clk:
        barrier()
        ev0.record()
        for _ in range(K):
            step_resident()
        ev1.record()
        barrier()
    ms = reduce_max(ev0.elapsed_time(ev1))
    launches = eng.launch_count() - l0 + (K if world > 1 else 0)
    res = {"value": world * B * K / (ms / 1e3), "ms_per_step": ms / K, "launches": int(launches), "clocks": clk.summary(),
           "T": T, "B": B}
    assert bool(torch.isfinite(out).all()), "non-finite samples in the timed batch"
    bytes_cs, flops_cs = eng.work(L)
    res["bytes_cs"], res["flops_cs"] = bytes_cs, flops_cs
    if not full:
        return res, eng, net

    # ---- clip 0 of the timed batch against a B=1 run of the same clip (rank 0) -------------
    if rank == 0:
        one = eng.sample(x_T[:1].contiguous(), noise[:, :1].contiguous(), coef, mel_spec=mel)
        d = (one - out[:1]).double()
        res["check"] = {"finite": True, "clip0_vs_b1_rel_l2": float(d.norm() / one.double().norm()),
                        "clip0_vs_b1_max_abs": float(d.abs().max()), "x0_std": float(out.std())}
        assert res["check"]["clip0_vs_b1_rel_l2"] < 1e-5, res["check"]
        eng.sample(x_T, noise, coef, mel_spec=mel, out=out)      # back to the B-clip workspace before the next timing
    barrier()

    # ---- end to end through sampling(): CPU draws + pinned staging + H2D inside the timed region, x_0 D2H ----
    x0_h = torch.empty((B, 1, L), pin_memory=True)

    def step_api():
        x0 = dwb.sampling(net, (B, 1, L), dh, condition=mel, verbose=False, out=out)
        x0_h.copy_(x0, non_blocking=True)

    for _ in range(2):
        step_api()
    barrier()
    ev0.record()
    for _ in range(K):
        step_api()
    ev1.record()
    barrier()
    res["e2e"] = world * B * K / (reduce_max(ev0.elapsed_time(ev1)) / 1e3)

    # ---- previous definition: draws made before the timer, H2D of the whole noise tensor inside ------------
    xd, nd = torch.empty_like(x_T), torch.empty_like(noise)

    def step_predrawn():
        xd.copy_(x_T_h, non_blocking=True)
        nd.copy_(noise_h, non_blocking=True)
        eng.sample(xd, nd, coef, mel_spec=mel, out=out)
        x0_h.copy_(out, non_blocking=True)

    step_predrawn()
    barrier()
    ev0.record()
    for _ in range(K):
        step_predrawn()
    ev1.record()
    barrier()
    res["e2e_predrawn"] = world * B * K / (reduce_max(ev0.elapsed_time(ev1)) / 1e3)
    res["h2d"] = (x_T.numel() + noise.numel()) * 4
    res["d2h"] = out.numel() * 4
    res["_x_T"], res["_mel"] = x_T, mel
    return res, eng, net


def roofline(name, res, eng, B, dev):
    import torch
    spec = CONFIGS[name]
    T = spec["T"]
    hbm, tf, src = peaks()
    step_s = res["ms_per_step"] / 1e3
    prof = eng.profile(res["_x_T"], torch.full((B,), float(T // 2), device=dev), res["_mel"], iters=3)
    tot = sum(v[0] for v in prof.values())
    cfg = spec["cfg"]
    kern = {}
    for kname, (kms, cnt) in prof.items():
        ent = {"kernel": kernel_label(kname, cfg, B), "ms_per_forward": round(kms, 4), "launches": cnt, "share": round(kms / tot, 4)}
        us = kms / cnt * 1e3
        if kname.startswith("fftconv_s") or kname.startswith("mix_s"):
            s = int(kname[-1])
            H, l = cfg["d_model"] * cfg["expand"] ** s, L // (4 ** s)
            nbytes = (2 if kname.startswith("fft") else 3) * 4.0 * H * l * B      # per launch
            ent["algorithmic_bytes_per_launch"] = nbytes
            ent["achieved_GBs"] = round(nbytes / (us * 1e-6) / 1e9, 1)
            ent["frac_of_hbm"] = round(ent["achieved_GBs"] / hbm, 4)
        if kname == "wave_block":
            C, S = cfg["res_channels"], cfg["skip_channels"]
            fl = (12.0 * C * C + 2.0 * C * C + 2.0 * C * S) * L * B             # per launch
            ent["algorithmic_flops_per_launch"] = fl
            ent["achieved_TFs"] = round(fl / (us * 1e-6) / 1e12, 1)
            ent["frac_of_tensor"] = round(ent["achieved_TFs"] / tf, 4)
            ent["mma_TFs_issued"] = round(3 * ent["achieved_TFs"], 1)
        kern[kname] = ent
    dom = max(kern, key=lambda k: kern[k]["share"])
    d = kern[dom]
    traffic = None                      # dram read+write bytes per launch from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get(f"{name}:{dom}", tj.get(dom, {}) if name == DEFAULT else {}).get(str(B))
    if spec["bound"] == "hbm":
        whole = res["bytes_cs"] * T * B / step_s / 1e9
        roof = {"bound": "hbm", "kernel": kernel_label(dom, cfg, B), "achieved": d.get("achieved_GBs"), "peak": hbm, "unit": "GB/s",
                "frac": d.get("frac_of_hbm"), "traffic": traffic, "peak_source": src,
                "algorithmic_bytes_per_launch": d.get("algorithmic_bytes_per_launch"),
                "whole_loop": {"achieved": round(whole, 1), "frac": round(whole / hbm, 4), "unit": "GB/s",
                               "algorithmic_bytes_per_clip_step": res["bytes_cs"], "flops_per_clip_step": res["flops_cs"],
                               "note": "SURVEY 8(d) bytes x T x B / step time"}}
    else:
        whole = res["flops_cs"] * T * B / step_s / 1e12
        roof = {"bound": "tensor", "kernel": kernel_label(dom, cfg, B), "achieved": d.get("achieved_TFs"), "peak": tf,
                "unit": "TFLOP/s", "frac": d.get("frac_of_tensor"), "traffic": traffic, "peak_source": src,
                "algorithmic_flops_per_launch": d.get("algorithmic_flops_per_launch"),
                "note": "fp32-equivalent work; every product is 3 bf16 MMAs (hi*hi + lo*hi + hi*lo, the parity mode), "
                        "so the tensor pipe issues 3x this figure and the single-pass bf16 peak caps frac at 1/3",
                "whole_loop": {"achieved": round(whole, 1), "frac": round(whole / tf, 4), "unit": "TFLOP/s",
                               "flops_per_clip_step": res["flops_cs"], "algorithmic_bytes_per_clip_step": res["bytes_cs"]}}
    roof["us_per_launch"] = round(d["ms_per_forward"] / d["launches"] * 1e3, 2)
    roof["share_of_step"] = d["share"]
    roof["kernels"] = kern
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU per step (default: the config's)")
    ap.add_argument("--config", default=DEFAULT, choices=list(CONFIGS))
    ap.add_argument("--impl", default="dwb", choices=["dwb", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner at communicator
    # creation) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path in the product"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    name = args.config
    spec = CONFIGS[name]
    B, K, W, T = args.batch or spec["batch"], args.steps, max(args.warmup, 3), spec["T"]

    res, eng, net = measure(name, B, K, W, dev, world, rank)
    line = None
    if rank == 0:
        roof = roofline(name, res, eng, B, dev)
        line = {"metric": metric_name(T), "value": round(res["value"], 4), "unit": "clips/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": round(res["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": spec["workload"], "name": name, "batch_per_gpu": B, "global_batch": B * world, "T": T, "L": L,
                           "parallelism": f"dp{world} (independent clips, one all_gather at sample collection)",
                           "l2": f"inputs larger than L2: {(T - 1) * B * L * 4 / 1e6:.0f} MB of noise + "
                                 f"{res['bytes_cs'] * B / 1e6:.0f} MB of activations streamed per diffusion step vs 126 MB L2"},
                "clocks": res["clocks"],
                "e2e": {"value": round(res["e2e"], 4), "unit": "clips/s", "h2d_bytes_per_step": res["h2d"],
                        "d2h_bytes_per_step": res["d2h"], "predrawn": round(res["e2e_predrawn"], 4),
                        "note": "sampling(net, size, dh): CPU-generator draws in the reference's order inside the timed region"},
                "gpu_launches": res["launches"], "check": res["check"], "roofline": roof}
    del eng, net, res
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and name == DEFAULT and not args.no_other_configs:
        hbm, tf, _ = peaks()
        others = {}
        for oname in ("unet_d128", "unet_d32_cond", "wnet_h256_d36"):
            ospec = CONFIGS[oname]
            r, e, n = measure(oname, ospec["batch"], 2, 1, dev, full=False)
            step_s = r["ms_per_step"] / 1e3
            ent = {"value": round(r["value"], 4), "unit": "clips/s", "batch": ospec["batch"], "T": ospec["T"],
                   "ms_per_step": round(r["ms_per_step"], 2), "steps": 2, "warmup": 1, "workload": ospec["workload"]}
            if ospec["bound"] == "hbm":
                ent["roofline_frac"] = round(r["bytes_cs"] * ospec["T"] * ospec["batch"] / step_s / 1e9 / hbm, 4)
                ent["bound"] = "hbm (whole loop: SURVEY 8(d) bytes / step time / measured HBM)"
            else:
                ent["roofline_frac"] = round(r["flops_cs"] * ospec["T"] * ospec["batch"] / step_s / 1e12 / tf, 4)
                ent["bound"] = "tensor (whole loop: fp32-equivalent flops / step time / measured bf16; 3-pass split caps it at 1/3)"
            others[oname] = ent
            del r, e, n
            torch.cuda.empty_cache()
        line["other_configs"] = others
        try:
            line["training_step"] = measure_training("wnet_h128_d30", 8, dev)
        except Exception as e:       # the training row must never cost the headline line
            line["training_step"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _, _ = cpu_arm(name, steps=2, warmup=1)
        line["cpu_baseline"] = cb
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if line:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
