import json,subprocess,os,sys
for tag,lib in (("A_skpre",""),("B_noskpre","/root/repo/diffwave_sashimi_b200/libdwb_noskpre.so"),("A_skpre2","")):
    env=dict(os.environ)
    if lib: env["DWB_LIB"]=lib
    r=subprocess.run([sys.executable,"bench.py","--steps","1","--warmup","1","--no-cpu-baseline","--no-other-configs"],capture_output=True,text=True,env=env)
    d=json.loads(r.stdout.strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(tag, d["value"], d["ms_per_step"], "mix_s0",k["mix_s0"]["ms_per_forward"],"mix_s1",k["mix_s1"]["ms_per_forward"],"mix_s2",k["mix_s2"]["ms_per_forward"], flush=True)
