"""C5-style sweep: atoms x replicas -> ms per evaluation, atoms*evals/s, parity vs the CPU oracle
on up to 256 randomly chosen replicas per point.  One GPU, or under torchrun: the R replicas of every point are sharded over the ranks
(contiguous blocks, sharding.shard_bounds), time = max over ranks, rank 0 prints.

    python tools/sweep_c5.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_c5.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import torch.distributed as dist  # noqa: E402

from gnnimplicitsolvent_b200 import sharding, synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402
from oracle import gnn_oracle as O  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
rows = []
for N in (10, 20, 50, 100, 200):
    mol = synth.synth_molecule(N, seed=N)
    for R in (64, 1024, 4096, 16384):
        if N * R > 1_700_000:
            continue
        R_total = R
        pos_all = synth.conformers(mol["pos"], R_total, 0.01, seed=N + R_total)
        lo, hi = sharding.shard_bounds(R_total, world, rank)
        R = hi - lo
        pos = np.ascontiguousarray(pos_all[lo:hi])
        sids = [r % 39 for r in range(lo, hi)]
        diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
        eng = Engine(mol["params"], sd, device=local)
        eng.set_replicas(R, sids, diel)
        p = torch.from_numpy(pos).cuda()
        e = torch.empty(R, device="cuda")
        f = torch.empty(R, N, 3, device="cuda")
        for _ in range(3):
            eng.energy_force(p, out_energy=e, out_force=f)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n_it = 10
        ev[0].record()
        for _ in range(n_it):
            eng.energy_force(p, out_energy=e, out_force=f)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n_it
        n_edges = eng.last_edge_count()
        if world > 1:
            t = torch.tensor([ms, float(n_edges)], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, n_edges = float(tmax[0]), int(t[1])
        e_err = f_err = None
        n_chk = 0
        if rank == 0:      # parity of randomly chosen replicas of this rank against the CPU oracle (256, fewer for large N)
            k = min(R, 256 if N <= 50 else (64 if N <= 100 else 16))
            idx = np.sort(np.random.default_rng(N * 1000 + R_total).choice(R, size=k, replace=False))
            o = O.evaluate(mol["params"], pos[idx], [sids[i] for i in idx], [diel[i] for i in idx], sd)
            ec, fc = e.cpu()[idx], f.cpu()[idx]
            e_err = float(((ec - o["energy_rep"]).abs() / o["energy_rep"].abs()).max())
            f_err = float(((fc - o["forces"]) ** 2).sum().sqrt() / (o["forces"] ** 2).sum().sqrt())
            n_chk = k
        rows.append({"atoms": N, "replicas": R_total, "gpus": world, "edges": n_edges, "ms_per_eval": ms,
                     "atoms_evals_per_s": N * R_total / (ms * 1e-3), "energy_max_rel": e_err, "force_rel_rms": f_err,
                     "replicas_checked": n_chk, "path": eng.describe()})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
        eng.close()
        del p, e, f
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
