import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
z = np.load("/root/repo/tests/golden/gnn_weights.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
mol = synth.synth_molecule(50, seed=50)
pos = torch.from_numpy(synth.conformers(mol["pos"], 4096, 0.01, seed=1)).cuda()
eng = Engine(mol["params"], sd); eng.set_replicas(4096, [0]*4096, [78.5]*4096)
for mode in (False, True, False, True):
    for _ in range(3): eng.energy_force(pos, tf32x1=mode)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): e, f = eng.energy_force(pos, tf32x1=mode)
    b.record(); torch.cuda.synchronize()
    print("tf32x1" if mode else "3xtf32", a.elapsed_time(b) / 10, "ms/eval")
