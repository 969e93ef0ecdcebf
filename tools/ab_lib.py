"""A/B two builds of libdwb.so on ONE box (interleaved runs of the whole bench): python tools/ab_lib.py path/to/other.so [label]
The alternative library is selected through DWB_LIB (diffwave_sashimi_b200/_lib.py)."""
import json, os, subprocess, sys
alt = os.path.abspath(sys.argv[1])
label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(alt)
for tag, lib in (("base", ""), (label, alt), ("base", ""), (label, alt)):
    env = dict(os.environ)
    if lib:
        env["DWB_LIB"] = lib
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-other-configs"],
                       capture_output=True, text=True, env=env)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(tag, d["value"], d["ms_per_step"], " ".join(f"{n} {k[n]['ms_per_forward']}" for n in k if n.startswith(("mix", "fft", "pool", "head"))), flush=True)
