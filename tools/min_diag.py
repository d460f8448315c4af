"""Development probe (GPU box): why does the line search stop?  Energy along the force direction at minimised points."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz")); sd = {k: torch.from_numpy(z[k]) for k in z.files}
N, R = 50, 64
mol = synth.synth_molecule(N, seed=50)
ff = synth.synth_forcefield(mol)
pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.02, seed=51)).cuda()
eng = Engine(mol["params"], sd); eng.set_forcefield(ff); eng.set_replicas(R, [0] * R, [78.5] * R)
p, e, st, rounds = eng.minimize(pos, grad_tol=10.0, max_iter=600)
nev, g = eng.minimize_stats()
print("status", np.bincount(st.cpu().numpy(), minlength=4).tolist(), "rounds", rounds, "grms median", float(g.median()))
def scan(label, **kw):
    e0, f0 = eng.energy_force(p, **kw)
    fn = f0.reshape(R, -1).norm(dim=1)                       # |F| per replica
    d = f0 / fn.reshape(R, 1, 1)
    rows = []
    for s in (1e-6, 1e-5, 1e-4, 3e-4, 1e-3, 3e-3):
        e1, _ = eng.energy_force(p + s * d, forces=False, **kw)
        e2, _ = eng.energy_force(p - s * d, forces=False, **kw)
        pred = -(fn * s)                                     # first-order change along +d (d = force direction: downhill)
        rows.append((s, float((e1 - e0).double().median()), float(pred.double().median()), float(((e1 - e2) / (2 * s) / (-fn)).double().median()),
                     float(((e1 - e2) / (2 * s) / (-fn)).double().min()), float(((e1 - e2) / (2 * s) / (-fn)).double().max())))
    print(label, "|F| median", float(fn.median()), "E median", float(e0.double().median()))
    for r_ in rows:
        print("   s=%.0e  dE(+s) median %.5f  predicted %.5f   central-difference slope / (-|F|): median %.4f min %.4f max %.4f" % r_)
scan("full (ff + GB + GNN)")
scan("vacuum force field only", vacuum_only=True)
scan("GB only + ff", gb_only=True)
eng2 = Engine(mol["params"], sd); eng2.set_replicas(R, [0] * R, [78.5] * R)
eng_saved = eng; eng = eng2
scan("GB + GNN without ff")
eng = eng_saved
# repeatability of the energy at the same point (fp32 noise floor)
ea, _ = eng.energy_force(p, forces=False); eb, _ = eng.energy_force(p + 1e-7, forces=False)
print("E(x+1e-7) - E(x): max abs", float((eb - ea).abs().max()))
