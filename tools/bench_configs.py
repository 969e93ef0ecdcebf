"""Per-forward time of the other BASELINE.json configs (parity-test cases, not bench lines): eager forwards,
CUDA events; clips/s = B / (T * s_per_forward).  python tools/bench_configs.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffwave_sashimi_b200 as dwb
from oracle.refshim import MODEL_CFGS      # config dictionaries only

CASES = [("unet_d128", 8, 200, None), ("wnet_h256_d36", 8, 200, None), ("unet_d32_cond", 16, 50, (1, 80, 63)), ("wnet_h128_d30", 8, 200, None)]
for name, B, T, melshape in CASES:
    cfg = dict(MODEL_CFGS[name])
    sd = dwb.init.seeded_state_dict(cfg, seed=0)
    net = dwb.construct_model(dict(cfg)); net.load_state_dict(sd); net = net.cuda().eval()
    x = torch.randn(B, 1, 16000, device="cuda"); t = torch.full((B, 1), 100.0, device="cuda")
    mel = torch.randn(*melshape, device="cuda") if melshape else None
    with torch.no_grad():
        for _ in range(3): net((x, t), mel_spec=mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): net((x, t), mel_spec=mel)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"config": name, "B": B, "T": T, "ms_per_forward": round(ms, 3), "clips_per_s_est": round(B / (T * ms / 1e3), 3)}))
    del net
    torch.cuda.empty_cache()
