"""Development probe (GPU box): batched minimiser statistics -- status histogram, rounds, final gradient RMS."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
ap = argparse.ArgumentParser()
ap.add_argument("--atoms", type=int, default=50); ap.add_argument("--replicas", type=int, default=512)
ap.add_argument("--tol", type=float, default=10.0); ap.add_argument("--iters", type=int, default=500)
ap.add_argument("--sigma", type=float, default=0.03); ap.add_argument("--ff", type=int, default=1)
ap.add_argument("--history", type=int, default=8); ap.add_argument("--net", type=int, default=0); ap.add_argument("--tag", default="")
a = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz")); sd = {k: torch.from_numpy(z[k]) for k in z.files}
mol = synth.synth_molecule(a.atoms, seed=a.atoms)
pos = torch.from_numpy(synth.conformers(mol["pos"], a.replicas, a.sigma, seed=3)).cuda()
eng = Engine(mol["params"], sd)
if a.net:
    iu = np.triu_indices(a.atoms, 1); bonds = np.stack(iu, axis=1).astype(np.int32)
    r0 = np.linalg.norm(mol["pos"][bonds[:, 0]] - mol["pos"][bonds[:, 1]], axis=1).astype(np.float32)
    eng.set_bonds(bonds, r0, np.full(len(bonds), 2.0e4, dtype=np.float32))
elif a.ff: eng.set_forcefield(synth.synth_forcefield(mol))
eng.set_replicas(a.replicas, [0] * a.replicas, [78.5] * a.replicas)
torch.cuda.synchronize(); t = time.perf_counter()
p, e, st, rounds = eng.minimize(pos, grad_tol=a.tol, max_iter=a.iters, history=a.history)
torch.cuda.synchronize(); dt = time.perf_counter() - t
e2, f = eng.energy_force(p)
g = f.reshape(a.replicas, -1).pow(2).mean(dim=1).sqrt().cpu().numpy()
st = st.cpu().numpy()
nev, gl = eng.minimize_stats(); nev = nev.cpu().numpy()
e0, _ = eng.energy_force(pos)
print(json.dumps({"tag": a.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("GNNIS_")}, "evals_pct": {q: int(np.percentile(nev, q)) for q in (5, 50, 95, 100)}, "N": a.atoms, "R": a.replicas, "tol": a.tol, "max_iter": a.iters, "rounds": rounds, "seconds": round(dt, 3),
  "status_hist": {int(k): int((st == k).sum()) for k in np.unique(st)},
  "grms_pct": {q: round(float(np.percentile(g, q)), 3) for q in (5, 50, 95, 100)},
  "grms_status0_max": round(float(g[st == 0].max()), 3) if (st == 0).any() else None,
  "grms_status1_pct": {q: round(float(np.percentile(g[st == 1], q)), 3) for q in (5, 50, 95)} if (st == 1).any() else None,
  "dE_mean": round(float((e2 - e0).mean()), 3)}))
