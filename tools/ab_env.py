"""A/B an environment switch on ONE box (interleaved runs of the whole bench): python tools/ab_env.py DWB_PDL=0 [more VAR=val ...]"""
import json, os, subprocess, sys
alt = dict(kv.split("=", 1) for kv in sys.argv[1:])
for tag, extra in (("base", {}), (str(alt), alt), ("base", {}), (str(alt), alt)):
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-other-configs"],
                       capture_output=True, text=True, env=dict(os.environ, **extra))
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print(tag, "FAILED", r.stderr[-600:])
        continue
    k = d["roofline"]["kernels"]
    print(tag, d["value"], d["ms_per_step"], d["check"]["clip0_vs_b1_rel_l2"], " ".join(f"{n} {k[n]['ms_per_forward']}" for n in k if n.startswith(("mix", "fft", "pool", "head"))), flush=True)
