import json, subprocess, sys
for B in (16, 32, 64, 128):
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-other-configs", "--batch", str(B)], capture_output=True, text=True)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print(B, d["value"], d["ms_per_step"], d["roofline"]["whole_loop"]["frac"], d["e2e"]["value"], flush=True)
