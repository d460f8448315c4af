"""C5-style sweep: atoms x replicas -> ms per evaluation, atoms*evals/s, parity vs the CPU oracle
on the first replicas.  One GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gnnimplicitsolvent_b200 import synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402
from oracle import gnn_oracle as O  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
rows = []
for N in (10, 20, 50, 100, 200):
    mol = synth.synth_molecule(N, seed=N)
    for R in (64, 1024, 4096, 16384):
        if N * R > 1_700_000:
            continue
        pos = synth.conformers(mol["pos"], R, 0.01, seed=N + R)
        sids = [r % 39 for r in range(R)]
        diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
        eng = Engine(mol["params"], sd)
        eng.set_replicas(R, sids, diel)
        p = torch.from_numpy(pos).cuda()
        e = torch.empty(R, device="cuda")
        f = torch.empty(R, N, 3, device="cuda")
        for _ in range(3):
            eng.energy_force(p, out_energy=e, out_force=f)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n_it = 10
        ev[0].record()
        for _ in range(n_it):
            eng.energy_force(p, out_energy=e, out_force=f)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n_it
        k = min(R, 4 if N >= 100 else 8)
        o = O.evaluate(mol["params"], pos[:k], sids[:k], diel[:k], sd)
        e_err = float(((e[:k].cpu() - o["energy_rep"]).abs() / o["energy_rep"].abs()).max())
        f_err = float(((f[:k].cpu() - o["forces"]) ** 2).sum().sqrt() / (o["forces"] ** 2).sum().sqrt())
        rows.append({"atoms": N, "replicas": R, "edges": eng.last_edge_count(), "ms_per_eval": ms,
                     "atoms_evals_per_s": N * R / (ms * 1e-3), "energy_max_rel": e_err, "force_rel_rms": f_err})
        print(json.dumps(rows[-1]), flush=True)
        eng.close()
        del p, e, f
        torch.cuda.empty_cache()
print(json.dumps({"sweep": rows}))
