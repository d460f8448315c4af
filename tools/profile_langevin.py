"""Small driver for ncu: a few LangevinMiddle steps of the C4-like system (N = 60, R = 1024, HBonds constraints)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gnnimplicitsolvent_b200 import synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
N, R = 60, 1024
mol = synth.synth_molecule(N, seed=60)
ff = synth.synth_forcefield(mol)
eng = Engine(mol["params"], sd)
eng.set_replicas(R, [0] * R, [78.5] * R)
eng.set_forcefield(ff)
hb = [i for i, (a, b) in enumerate(ff["bonds"]["idx"]) if mol["elements"][a].startswith("H") or mol["elements"][b].startswith("H")]
eng.set_constraints(ff["bonds"]["idx"][hb], ff["bonds"]["r0"][hb])
p = torch.from_numpy(synth.conformers(mol["pos"], R, 0.0, seed=61)).cuda().contiguous()
v = torch.zeros_like(p)
m = np.asarray(mol["masses"], dtype=np.float32)
eng.langevin(p, v, m, int(sys.argv[1]) if len(sys.argv) > 1 else 3, temperature=300.0, friction=5.0, dt=0.002, seed=1)
torch.cuda.synchronize()
print("finite", bool(torch.isfinite(p).all()))
