"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share (markdown)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    k = r[c["Kernel Name"]].split("(")[0].replace("void ", "").replace("dwb::", "")
    v = float(r[c["Metric Value"]].replace(",", ""))
    u = r[c["Metric Unit"]]
    v = v / 1000 if u.startswith("ns") or u == "nsecond" else (v * 1000 if u.startswith("ms") else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total µs | share |\n|---|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | 100% |")
