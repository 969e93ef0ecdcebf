"""Summarise an `ncu --page source --csv` dump: total stall-reason samples and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:          # first kernel instance only (the dump repeats the header per instance)
    if len(r) != len(hdr) or r[0] == "Address":
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:90], "| SASS lines", len(body), "| samples", tot,
      "| warp instr", sum(int(r[col["Instructions Executed"]] or 0) for r in body))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((sum(int(r[col[h]] or 0) for r in body), h) for h in reasons), reverse=True)
print("stalls:", ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for v, h in agg[:9]))
top = sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]
for r in top:
    rs = sorted(((int(r[col[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{100 * int(r[col['# Samples']]) / tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} {rs}")
ops = {}
for r in body:
    toks = r[col["Source"]].strip().split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[col["Instructions Executed"]] or 0)
ti = sum(ops.values())
print("opcodes:", ", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
