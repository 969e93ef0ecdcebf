"""Race libdwb's Cauchy kernels against the UNMODIFIED reference extension rebuilt for sm_100a (baseline/_ref/cauchy,
built by oracle/build_ref_cauchy.py) on the same tensors, and time the fused S4 kernel generation against the
reference recipe (Cauchy kernel + Woodbury in torch).  CUDA events, median of `--repeat` launches after warm-up.

    python tools/bench_cauchy.py [--repeat 30] [--json gpurun_out/cauchy.json]

Shapes: the reference's own benchmark (extensions/cauchy/benchmark_cauchy.py:29-42: batch 1024, N = 64 -> half 32,
L = 16384) and the three calls one S4 layer of unet d64 makes at plan time (models/s4.py:758: v (6H, 32), z (l/2+1)).
Algorithmic bytes: reads batch*N*16 + L*8, writes batch*L*8.
"""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diffwave_sashimi_b200 as dwb  # noqa: E402


def timed(fn, repeat):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(repeat):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def load_reference():
    d = os.path.join(ROOT, "baseline", "_ref", "cauchy")
    if not os.path.exists(os.path.join(d, "cauchy_mult.so")):
        return None
    sys.path.insert(0, d)
    try:
        return importlib.import_module("cauchy_mult")
    except Exception as e:          # noqa: BLE001
        print("reference extension not loadable:", e, file=sys.stderr)
        return None
    finally:
        sys.path.remove(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeat", type=int, default=30)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ref = load_reference()
    dev = torch.device("cuda")
    hbm = 6545.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = json.load(open(p))["hbm_gbs"]
    out = {"reference_extension": "baseline/_ref/cauchy/cauchy_mult.so (unmodified cauchy.cpp + cauchy_cuda.cu, sm_100a, "
                                  "-O3 --use_fast_math)" if ref else None, "hbm_peak_GBs": hbm, "cases": []}
    g = torch.Generator(device="cuda").manual_seed(2357)
    shapes = [("benchmark_cauchy.py", 1024, 32, 16384), ("d64 top stage (6H=384, l=16000)", 384, 32, 8001),
              ("d64 mid stage (6H=768, l=4000)", 768, 32, 2001), ("d64 centre (6H=1536, l=1000)", 1536, 32, 501)]
    for name, batch, N, L in shapes:
        v = torch.randn(batch, N, dtype=torch.complex64, device=dev, generator=g)
        w = torch.randn(batch, N, dtype=torch.complex64, device=dev, generator=g)
        z = torch.exp(1j * torch.randn(L, device=dev, generator=g)).to(torch.complex64)
        dout = torch.randn(batch, L, dtype=torch.complex64, device=dev, generator=g)
        ent = {"case": name, "batch": batch, "N_half": N, "L": L}
        byts = batch * N * 16 + L * 8 + batch * L * 8
        ours = timed(lambda: dwb.ops.cauchy_mult_sym_fwd(v, z, w), args.repeat)
        ent["sym_fwd"] = {"dwb_us": round(ours, 2), "algorithmic_bytes": byts, "dwb_GBs": round(byts / ours / 1e3, 1),
                          "dwb_frac_of_hbm": round(byts / ours / 1e3 / hbm, 4)}
        ours_b = timed(lambda: dwb.ops.cauchy_mult_sym_bwd(v, z, w, dout), args.repeat)
        ent["sym_bwd"] = {"dwb_us": round(ours_b, 2)}
        if ref:
            r = timed(lambda: ref.cauchy_mult_sym_fwd(v, z, w), args.repeat)
            a, b = dwb.ops.cauchy_mult_sym_fwd(v, z, w), ref.cauchy_mult_sym_fwd(v, z, w)
            ent["sym_fwd"].update(reference_us=round(r, 2), speedup=round(r / ours, 3),
                                  max_rel_diff=float((a - b).abs().max() / b.abs().max()))
            rb = timed(lambda: ref.cauchy_mult_sym_bwd(v, z, w, dout), args.repeat)
            (dv, dw_), (rv, rw) = dwb.ops.cauchy_mult_sym_bwd(v, z, w, dout), ref.cauchy_mult_sym_bwd(v, z, w, dout)
            ent["sym_bwd"].update(reference_us=round(rb, 2), speedup=round(rb / ours_b, 3),
                                  max_rel_diff_dv=float((dv - rv).abs().max() / rv.abs().max()),
                                  max_rel_diff_dw=float((dw_ - rw).abs().max() / rw.abs().max()))
        out["cases"].append(ent)

    # fused kernel generation of one S4 layer (Cauchy + Woodbury + irfft) vs the reference recipe with its own kernel
    for H, l in ((64, 16000), (128, 4000), (256, 1000)):
        sd = dwb.init.s4_layer_params(H, generator=torch.Generator().manual_seed(1))
        P = {k.split(".")[-1]: t.to(dev) for k, t in sd.items()}
        C = dwb.engine.setup_C(P["C"], P["B"], P["P"], P["inv_w_real"], P["w_imag"], P["log_dt"], l)
        om = dwb.engine.reference_nodes(l).to(dev)
        fused = timed(lambda: dwb.ops.s4_kernel_gen(C, P["B"], P["P"], P["inv_w_real"], P["w_imag"], P["log_dt"], l, omega=om),
                      max(5, args.repeat // 3))
        ent = {"case": f"S4 kernel generation H={H} l={l} (models/s4.py:674-807)", "dwb_fused_fp64_us": round(fused, 1)}
        if ref:
            Cc, Bc, Pc = (torch.view_as_complex(t.contiguous()) for t in (C, P["B"], P["P"]))
            wq = torch.complex(-torch.exp(P["inv_w_real"]), P["w_imag"])
            dt = torch.exp(P["log_dt"])
            z = 2 * (1 - om) / (1 + om)

            def recipe():
                Bt, Ct = torch.cat([Bc, Pc], 0), torch.cat([Cc, Pc.conj()], 0)          # (2,H,N), (3,H,N)
                v = (Bt.unsqueeze(1) * Ct.unsqueeze(0)).reshape(-1, Cc.shape[-1])      # (6H, N)
                ww = (wq * dt[:, None]).repeat(6, 1)
                r = ref.cauchy_mult_sym_fwd(v.contiguous(), z.contiguous(), ww.contiguous()).view(2, 3, H, -1) * dt[None, None, :, None]
                kf = r[:1, :2] - r[:1, 2:] * r[1:, :2] / (1 + r[1:, 2:])
                kf = kf * 2 / (1 + om)
                return torch.fft.irfft(kf, n=l)

            ent["reference_recipe_fp32_us"] = round(timed(recipe, max(5, args.repeat // 3)), 1)
            ent["speedup"] = round(ent["reference_recipe_fp32_us"] / fused, 3)
            ent["note"] = "reference: its Cauchy kernel + Woodbury/irfft in torch (complex64, cuFFT); dwb: one fused fp64 kernel + fp64 direct irDFT"
        out["cases"].append(ent)
    s = json.dumps(out, indent=1)
    print(s)
    if args.json:
        open(args.json, "w").write(s)


if __name__ == "__main__":
    main()
