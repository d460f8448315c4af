"""Small driver for ncu: two evaluations of the C2 workload at a reduced replica count (same N, same weights)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gnnimplicitsolvent_b200 import synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=1024)
ap.add_argument("--atoms", type=int, default=50)
ap.add_argument("--evals", type=int, default=2)
ap.add_argument("--fp16", action="store_true", help="half-precision mode of the edge MLPs (GNNIS_FLAG_MLP_FP16)")
ap.add_argument("--bf16", action="store_true")
ap.add_argument("--chain", action="store_true", help="extended random-walk geometry, parameter rows repeated (large molecules)")
a = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
if a.chain:
    mol = {"params": np.resize(synth.synth_molecule(50, seed=50)["params"], (a.atoms, 7))}
    rng = np.random.default_rng(3)
    d = rng.normal(size=(a.replicas, a.atoms, 3))
    d[..., 0] += 1.5
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    pos = torch.from_numpy((np.cumsum(0.12 * d, axis=1) + rng.normal(scale=0.05, size=d.shape)).astype(np.float32)).cuda()
else:
    mol = synth.synth_molecule(a.atoms, seed=a.atoms)
    pos = torch.from_numpy(synth.conformers(mol["pos"], a.replicas, 0.01, seed=1)).cuda()
eng = Engine(mol["params"], sd)
eng.set_replicas(a.replicas, [0] * a.replicas, [78.5] * a.replicas)
for _ in range(a.evals):
    e, f = eng.energy_force(pos, fp16=a.fp16, bf16=a.bf16)
torch.cuda.synchronize()
print("E sum", float(e.double().sum()), "edges", eng.last_edge_count(), "launches", eng.launch_count())
