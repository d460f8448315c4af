"""Per-kernel SASS opcode histogram of libgnnis.so (cuobjdump -sass): the tensor-core / TMA / MUFU evidence asked for in
/opt/skills/guides/B200_PROFILING.md.  usage: python tools/sass_histogram.py [lib] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gnnimplicitsolvent_b200", "libgnnis.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU.EX2", "MUFU.RCP", "MUFU.TANH",
        "MUFU.RSQ", "MUFU.LG2", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "PRMT", "LDS", "STS", "LDG", "STG", "LDL", "STL",
        "SHFL", "BAR", "ATOMS", "RED", "ATOMG", "F2FP"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
func, hist, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        func = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        func = re.sub(r"\(.*", "", func).replace("gnnis::", "")
        hist[func] = collections.Counter()
        total[func] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and func:
        op = m.group(1)
        total[func] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k.startswith("MUFU") and op.startswith(k)):
                hist[func][k] += 1
                break
print(f"# SASS opcode counts per kernel of {os.path.basename(lib)} (static instruction counts, cuobjdump -sass, sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine, non-tensor),")
print("# SYNCS = mbarrier ops, FFMA2 / FMUL2 / FADD2 = packed fp32 pairs; there is no UTMALDG / UTMASTG (no tensor-map TMA)")
for f in sorted(hist, key=lambda x: -total[x]):
    if total[f] < 40:
        continue
    items = " ".join(f"{k}={v}" for k, v in hist[f].items() if v)
    print(f"{f}  [{total[f]} instructions]  {items}")
