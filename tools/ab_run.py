"""A/B helper (development only, GPU box): for every _ab/*.so, stage times + parity of the C2 workload.
usage: python tools/ab_run.py [--replicas 4096] [names...]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from gnnimplicitsolvent_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
R = int(sys.argv[2]); N = int(sys.argv[3])
z = np.load(os.path.join(%r, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
mol = synth.synth_molecule(N, seed=N)
pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=1)).cuda()
eng = Engine(mol["params"], sd)
eng.set_replicas(R, [0] * R, [78.5] * R)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    e, f = eng.energy_force(pos)
acc = None
n = 10
for _ in range(n):
    flush.fill_(1)
    st = eng.profile_eval(pos)
    acc = st if acc is None else {k: acc[k] + v for k, v in st.items()}
torch.cuda.synchronize()
e, f = eng.energy_force(pos)
out = {k: round(v / n, 4) for k, v in acc.items()}
out["total"] = round(sum(acc.values()) / n, 4)
out["E"] = float(e.double().sum()); out["F2"] = float((f.double() ** 2).sum())
print(json.dumps(out))
''' % (ROOT, ROOT)
args = sys.argv[1:]
R, N = 4096, 50
if "--replicas" in args:
    i = args.index("--replicas"); R = int(args[i + 1]); del args[i:i + 2]
if "--atoms" in args:
    i = args.index("--atoms"); N = int(args[i + 1]); del args[i:i + 2]
libs = sorted(glob.glob(os.path.join(ROOT, "_ab", "*.so")))
if args:
    libs = [l for l in libs if os.path.basename(l)[:-3] in args]
for rep in range(2):
    for lib in libs:
        r = subprocess.run([sys.executable, "-c", CHILD, lib, str(R), str(N)], capture_output=True, text=True)
        print(os.path.basename(lib), r.stdout.strip() or r.stderr[-800:], flush=True)
