"""Development tool (GPU box): stage times of one workload plus parity against the CPU oracle on a random sample of
replicas, printed as one JSON line.  Environment switches of libgnnis (GNNIS_STASH, GNNIS_TC, ...) are read by the
library itself, so variants are compared by running this script once per setting.

usage: python tools/stage_times.py [--replicas 4096] [--atoms 50] [--check 32] [--flags 0] [--lib path/to/libgnnis.so]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=4096)
ap.add_argument("--atoms", type=int, default=50)
ap.add_argument("--check", type=int, default=32)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--lib", default=None)
ap.add_argument("--tag", default="")
a = ap.parse_args()
from gnnimplicitsolvent_b200 import _lib  # noqa: E402

if a.lib:
    _lib.LIB_PATH = a.lib
from gnnimplicitsolvent_b200 import synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402
from oracle import gnn_oracle as O  # noqa: E402

R, N = a.replicas, a.atoms
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
mol = synth.synth_molecule(N, seed=N)
pos_h = synth.conformers(mol["pos"], R, 0.01, seed=1)
pos = torch.from_numpy(pos_h).cuda()
eng = Engine(mol["params"], sd)
eng.set_replicas(R, [0] * R, [78.5] * R)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
kw = {"extra_flags": a.flags} if a.flags else {}
for _ in range(3):
    e, f = eng.energy_force(pos, **kw)
torch.cuda.synchronize()
acc = None
for _ in range(a.iters):
    flush.fill_(1)
    st = eng.profile_eval(pos, **kw)
    acc = st if acc is None else {k: acc[k] + v for k, v in st.items()}
torch.cuda.synchronize()
# whole evaluation through the graph-replay path
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(a.iters):
    flush.fill_(1)
    ev0.record()
    e, f = eng.energy_force(pos, **kw)
    ev1.record()
    torch.cuda.synchronize()
    tot += ev0.elapsed_time(ev1)
out = {"tag": a.tag, "N": N, "R": R, "env": {k: v for k, v in os.environ.items() if k.startswith("GNNIS_")}}
out.update({k: round(v / a.iters, 4) for k, v in acc.items()})
out["stages_total"] = round(sum(acc.values()) / a.iters, 4)
out["eval_ms"] = round(tot / a.iters, 4)
out["edges"] = eng.last_edge_count()
# parity on a random sample of replicas
if a.check > 0:
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(R, size=min(a.check, R), replace=False))
    o = O.evaluate(mol["params"], pos_h[idx], [0] * len(idx), [78.5] * len(idx), sd)
    ec, fc = e.cpu()[idx], f.cpu()[idx]
    out["E_relerr"] = float((ec - o["energy_rep"]).abs().max() / o["energy_rep"].abs().max())
    out["F_relrms"] = float(((fc - o["forces"]) ** 2).sum().sqrt() / (o["forces"] ** 2).sum().sqrt())
    per = ((fc - o["forces"]) ** 2).sum(dim=(1, 2)).sqrt() / (o["forces"] ** 2).sum(dim=(1, 2)).sqrt()
    out["F_relrms_worst_replica"] = float(per.max())
out["finite"] = bool(torch.isfinite(e).all() and torch.isfinite(f).all())
print(json.dumps(out), flush=True)
