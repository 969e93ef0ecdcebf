"""Top source lines by stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
fname, out, hdr = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        c = {h: i for i, h in enumerate(hdr)}
        smp = int(r[c["# Samples"]]) if r[c["# Samples"]].isdigit() else 0
        ins = int(r[c["Instructions Executed"]]) if r[c["Instructions Executed"]].isdigit() else 0
        out.append((smp, ins, fname, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out)
toti = sum(o[1] for o in out)
print("total samples", tot, "warp instr", toti)
for smp, ins, f, ln, src in sorted(out, reverse=True)[:n]:
    print(f"{100*smp/tot:5.1f}% smp {100*ins/toti:5.1f}% ins  {f}:{ln}  {src}")
