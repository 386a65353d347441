"""Top stall locations of one kernel from an .ncu-rep source page:  python scripts/ncu_source_top.py rep.ncu-rep <kernel regex> [N]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)   # first matching launch only
h = rows[hi]
d = [r for r in rows[hi + 1:end] if len(r) == len(h) and r[0] != "Address"]
si, src = h.index("# Samples"), h.index("Source")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
num = lambda v: int(float(v)) if v not in ("", None) else 0
tot = sum(num(r[si]) for r in d) or 1
print("total samples", tot, "instructions", len(d))
for r in sorted(d, key=lambda r: -num(r[si]))[:N]:
    st = sorted(((num(r[i]), h[i][6:]) for i in stalls), reverse=True)[:2]
    print(f"{num(r[si]):6d} {100 * num(r[si]) / tot:5.1f}%  {r[src][:100]:100s} {st}")
agg = {h[i][6:]: sum(num(r[i]) for r in d) for i in stalls}
print("by reason:", [(k, f"{100 * v / tot:.1f}%") for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
