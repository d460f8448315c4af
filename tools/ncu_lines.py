"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export: stall samples and
executed warp instructions per line of the .cu file (lines carrying at least `frac` of all samples)."""
import csv
import sys

path = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
rows = list(csv.reader(open(path)))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] != "":          # a source line row (aggregated over its SASS)
        lines.append(r)
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
tot = sum(num(r[ix["# Samples"]]) for r in lines)
toti = sum(num(r[ix["Instructions Executed"]]) for r in lines)
print("samples", tot, "warp instructions", toti)
keys = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected",
        "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch"]
agg = {k: sum(num(r[ix[k]]) for r in lines) for k in keys}
print({k: f"{100 * v / tot:.1f}%" for k, v in agg.items()})
for r in lines:
    s = num(r[ix["# Samples"]])
    n = num(r[ix["Instructions Executed"]])
    if s >= tot * frac or n >= toti * frac:
        st = {k[6:]: num(r[ix[k]]) for k in keys if num(r[ix[k]]) > 0.2 * s and s > 0}
        print(f"{r[0]:>5} {100 * s / tot:5.1f}% smp {100 * n / toti:5.1f}% ins  {r[1][:90]}  {st}")
