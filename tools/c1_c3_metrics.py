"""The two BASELINE.json configurations that bench.py does not time:

C1  N = 13 (COCCO + H shape), R = 128, DMSO: energy+force time and minimise-to-convergence time on one GPU, next to the
    CPU oracle (the reference's arithmetic) on the host cores for the same batch.
C3  N = 50, R = 39 x 512 = 19 968 replicas, solvent id = replica mod 39 (helper_functions.py:738-743), sharded over the
    ranks of torchrun (contiguous blocks, one all_gather of the energies); parity vs the oracle on a few replicas.

    python tools/c1_c3_metrics.py                         # C1 and C3 on 1 GPU
    torchrun --nproc-per-node 8 tools/c1_c3_metrics.py     # C3 sharded over 8 GPUs (rank 0 also runs C1)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gnnimplicitsolvent_b200 import sharding, synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
out = {}


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


if rank == 0:
    # ---------------------------------------------------------------- C1
    from oracle import gnn_oracle as O
    N, R = 13, 128
    mol = synth.synth_molecule(N, seed=13)
    pos = synth.conformers(mol["pos"], R, 0.02, seed=1013)
    eng = Engine(mol["params"], sd, device=local)
    eng.set_replicas(R, [3] * R, [47.24] * R)
    p = torch.from_numpy(pos).cuda()
    ms = timed(lambda: eng.energy_force(p), 50)
    t0 = time.perf_counter()
    o = O.evaluate(mol["params"], pos, [3] * R, [47.24] * R, sd)
    t_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    o = O.evaluate(mol["params"], pos, [3] * R, [47.24] * R, sd)
    t_cpu = min(t_cpu, time.perf_counter() - t0)
    e, f = eng.energy_force(p)
    eng.set_forcefield(synth.synth_forcefield(mol))
    eng.minimize(p, grad_tol=10.0, max_iter=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pm, em, status, rounds = eng.minimize(p, grad_tol=10.0, max_iter=500)
    torch.cuda.synchronize()
    t_min = time.perf_counter() - t0
    out["c1"] = {"atoms": N, "replicas": R, "solvent": "DMSO (id 3, eps 47.24)", "gpu_ms_per_eval": ms,
                 "gpu_atoms_evals_per_s": N * R / (ms * 1e-3), "cpu_oracle_ms_per_eval": 1e3 * t_cpu,
                 "cpu_cores": torch.get_num_threads(), "speedup_vs_cpu": 1e3 * t_cpu / ms,
                 "energy_max_rel": float((e.cpu() - o["energy_rep"]).abs().max() / o["energy_rep"].abs().max()),
                 "minimize": {"potential": "GNN + synthetic vacuum force field", "grad_tol_kJ_mol_nm": 10.0,
                              "rounds": rounds, "seconds": t_min,
                              "status_counts": {str(k): int((status == k).sum()) for k in range(4)},
                              "final_grad_rms_median": float(eng.minimize_stats()[1].median()),
                              "evaluations_per_replica_median": float(eng.minimize_stats()[0].float().median()),
                              "s_per_conformer": t_min / R}}
    eng.close()

# -------------------------------------------------------------------- C3
N, S, PER = 50, 39, 512
R = S * PER
lo, hi = sharding.shard_bounds(R, world, rank)
mol = synth.synth_molecule(N, seed=50)
sids_all = sharding.interleaved_solvents(R, S)
diel_all = [synth.SOLVENT_DIELECTRICS[s] for s in sids_all]
pos_local = synth.conformers(mol["pos"], R, 0.01, seed=50039)[lo:hi]     # the same global batch on every rank
eng = Engine(mol["params"], sd, device=local)
eng.set_replicas(hi - lo, sids_all[lo:hi], diel_all[lo:hi])
p = torch.from_numpy(np.ascontiguousarray(pos_local)).cuda()
if world > 1:
    dist.barrier()
ms = timed(lambda: eng.energy_force(p), 10)
t = torch.tensor([ms], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
e, f = eng.energy_force(p)
e_all = sharding.gather_energies(e, R, world, rank)
if rank == 0:
    from oracle import gnn_oracle as O
    k = 6
    o = O.evaluate(mol["params"], pos_local[:k], sids_all[:k], diel_all[:k], sd)
    out["c3"] = {"atoms": N, "replicas": R, "solvents": S, "gpus": world, "replicas_per_gpu": hi - lo,
                 "ms_per_eval_max_over_ranks": float(t.item()), "atoms_evals_per_s": N * R / (float(t.item()) * 1e-3),
                 "energies_gathered": int(e_all.numel()), "energy_checksum": float(e_all.double().sum()),
                 "energy_max_rel_first6": float((e[:k].cpu() - o["energy_rep"]).abs().max() / o["energy_rep"].abs().max()),
                 "force_rel_rms_first6": float(((f[:k].cpu() - o["forces"]) ** 2).sum().sqrt() / (o["forces"] ** 2).sum().sqrt())}
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
