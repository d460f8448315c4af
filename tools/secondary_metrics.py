"""Secondary metrics of SURVEY.md 8(d): batched minimisation throughput (C2) and Langevin ns/day (C4-like) on the
full Hamiltonian: GNN implicit solvent + synthetic vacuum force field (synth.synth_forcefield), HBonds constraints."""
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
    return mol, pos, synth.synth_forcefield(mol)


z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
out = {}

# C2: N = 50, R = 4096, water: L-BFGS to gradient RMS < 10 kJ/mol/nm
mol, pos, ff = system(50, 4096, 50, 0.02)
eng = Engine(mol["params"], sd)
eng.set_replicas(4096, [0] * 4096, [78.5] * 4096)
eng.set_forcefield(ff)
p0 = torch.from_numpy(pos).cuda()
eng.minimize(p0[:], grad_tol=10.0, max_iter=3)      # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
p1, e1, status, rounds = eng.minimize(p0, grad_tol=10.0, max_iter=300)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
n_ev, grms = eng.minimize_stats()
n_ev, grms, st_ = n_ev.cpu().numpy(), grms.cpu().numpy(), status.cpu().numpy()
pct = lambda a, q: float(np.percentile(a, q))
out["minimize_c2"] = {"replicas": 4096, "atoms": 50, "grad_tol": 10.0, "max_iter": 300, "rounds": rounds, "seconds": dt,
                      "status_counts": {"0_below_threshold": int((st_ == 0).sum()), "1_iteration_limit": int((st_ == 1).sum()),
                                        "2_non_finite": int((st_ == 2).sum()), "3_line_search_exhausted": int((st_ == 3).sum())},
                      "evaluations_per_replica": {"median": pct(n_ev, 50), "p95": pct(n_ev, 95), "max": int(n_ev.max())},
                      "final_grad_rms": {"p5": pct(grms, 5), "median": pct(grms, 50), "p95": pct(grms, 95), "max": float(grms.max())},
                      "energy_drop_mean_kj_mol": float((eng.energy_force(p0)[0] - e1).mean()),
                      "conformers_per_s": 4096 / dt, "ms_per_round": 1e3 * dt / max(rounds, 1)}
# the same batch with an elastic network instead of the synthetic force field: every conformer has ONE well-defined basin
eng.set_forcefield(None)
iu = np.triu_indices(50, 1)
bonds = np.stack(iu, axis=1).astype(np.int32)
r0 = np.linalg.norm(mol["pos"][bonds[:, 0]] - mol["pos"][bonds[:, 1]], axis=1).astype(np.float32)
eng.set_bonds(bonds, r0, np.full(len(bonds), 2.0e4, dtype=np.float32))
torch.cuda.synchronize()
t0 = time.perf_counter()
p2, e2, status2, rounds2 = eng.minimize(p0, grad_tol=10.0, max_iter=300)
torch.cuda.synchronize()
dt2 = time.perf_counter() - t0
n_ev2, grms2 = eng.minimize_stats()
n_ev2, grms2, st2 = n_ev2.cpu().numpy(), grms2.cpu().numpy(), status2.cpu().numpy()
out["minimize_c2_elastic_network"] = {"rounds": rounds2, "seconds": dt2, "status_counts": {str(k): int((st2 == k).sum()) for k in range(4)},
                                      "evaluations_per_replica": {"median": pct(n_ev2, 50), "p95": pct(n_ev2, 95), "max": int(n_ev2.max())},
                                      "final_grad_rms": {"median": pct(grms2, 50), "p95": pct(grms2, 95), "max": float(grms2.max())},
                                      "conformers_per_s": 4096 / dt2}
eng.close()

# C4-like: N = 60, R = 1024, 300 K, 1/ps, dt = 2 fs with HBonds constraints
mol, pos, ff = system(60, 1024, 60, 0.0)
eng = Engine(mol["params"], sd)
eng.set_replicas(1024, [0] * 1024, [78.5] * 1024)
eng.set_forcefield(ff)
hb = [i for i, (a, b) in enumerate(ff["bonds"]["idx"]) if mol["elements"][a].startswith("H") or mol["elements"][b].startswith("H")]
eng.set_constraints(ff["bonds"]["idx"][hb], ff["bonds"]["r0"][hb])
p, _, st_min, _ = eng.minimize(torch.from_numpy(pos).cuda(), grad_tol=20.0, max_iter=300)   # relax the synthetic start
p = p.contiguous()
v = torch.zeros_like(p)
m = np.asarray(mol["masses"], dtype=np.float32)
eng.langevin(p, v, m, 20, temperature=300.0, friction=5.0, dt=0.002, seed=1)
torch.cuda.synchronize()
steps = 1000
t0 = time.perf_counter()
eng.langevin(p, v, m, steps, temperature=300.0, friction=5.0, dt=0.002, seed=2, first_step=20)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t_step = dt / steps
out["langevin_c4"] = {"replicas": 1024, "atoms": 60, "steps": steps, "ms_per_step": 1e3 * t_step,
                      "ns_per_day_per_replica_at_2fs": 86400.0 * 2e-6 / t_step,
                      "aggregate_ns_per_day_at_2fs": 1024 * 86400.0 * 2e-6 / t_step,
                      "finite": bool(torch.isfinite(p).all()),
                      "kinetic_T_over_300K": float((torch.from_numpy(m).cuda().reshape(1, -1, 1) * v * v).sum() / 1024
                                                   / ((3 * 60 - len(hb)) * 0.0083144626 * 300.0))}
print(json.dumps(out))
