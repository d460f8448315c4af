"""Secondary metrics of SURVEY.md 8(d): batched minimisation throughput (C2) and Langevin ns/day (C4-like).
GNN potential + an elastic-network stand-in for the vacuum force field (SURVEY.md N1 is not built yet)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gnnimplicitsolvent_b200 import synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402


def system(N, R, seed, sigma):
    mol = synth.synth_molecule(N, seed=seed)
    pos = synth.conformers(mol["pos"], R, sigma, seed=seed + 1)
    iu = np.triu_indices(N, 1)
    bonds = np.stack(iu, axis=1).astype(np.int32)
    r0 = np.linalg.norm(mol["pos"][bonds[:, 0]] - mol["pos"][bonds[:, 1]], axis=1).astype(np.float32)
    return mol, pos, bonds, r0, np.full(len(bonds), 2.0e4, dtype=np.float32)


z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
out = {}

# C2: N = 50, R = 4096, water: L-BFGS to gradient RMS < 10 kJ/mol/nm
mol, pos, bonds, r0, k = system(50, 4096, 50, 0.03)
eng = Engine(mol["params"], sd)
eng.set_replicas(4096, [0] * 4096, [78.5] * 4096)
eng.set_bonds(bonds, r0, k)
p0 = torch.from_numpy(pos).cuda()
eng.minimize(p0[:], grad_tol=10.0, max_iter=3)      # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
p1, e1, status, rounds = eng.minimize(p0, grad_tol=10.0, max_iter=300)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["minimize_c2"] = {"replicas": 4096, "atoms": 50, "rounds": rounds, "seconds": dt,
                      "converged": int((status == 0).sum()), "conformers_per_s": 4096 / dt,
                      "ms_per_round": 1e3 * dt / max(rounds, 1)}
eng.close()

# C4-like: N = 60, R = 1024, 300 K, 1/ps, dt = 2 fs (no constraints: 0.5 fs would be needed for real hydrogens;
# the timing does not depend on dt)
mol, pos, bonds, r0, k = system(60, 1024, 60, 0.002)
eng = Engine(mol["params"], sd)
eng.set_replicas(1024, [0] * 1024, [78.5] * 1024)
eng.set_bonds(bonds, r0, k)
p = torch.from_numpy(pos).cuda().contiguous()
v = torch.zeros_like(p)
m = np.asarray(mol["masses"], dtype=np.float32)
eng.langevin(p, v, m, 20, temperature=300.0, friction=1.0, dt=0.0005, seed=1)
torch.cuda.synchronize()
steps = 500
t0 = time.perf_counter()
eng.langevin(p, v, m, steps, temperature=300.0, friction=1.0, dt=0.0005, seed=2, first_step=20)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t_step = dt / steps
out["langevin_c4"] = {"replicas": 1024, "atoms": 60, "steps": steps, "ms_per_step": 1e3 * t_step,
                      "ns_per_day_per_replica_at_2fs": 86400.0 * 2e-6 / t_step,
                      "aggregate_ns_per_day_at_2fs": 1024 * 86400.0 * 2e-6 / t_step,
                      "finite": bool(torch.isfinite(p).all())}
print(json.dumps(out))
