"""Probe: evaluation time at small sizes under different GNNIS_TC masks (development)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
for N, R in [(50, 1), (50, 8), (13, 128), (100, 4), (50, 64), (50, 256)]:
    mol = synth.synth_molecule(N, seed=N)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=1)).cuda()
    out = []
    for mask in ("7", "3"):
        os.environ["GNNIS_TC"] = mask
        eng = Engine(mol["params"], sd)
        eng.set_replicas(R, [0] * R, [78.5] * R)
        e = torch.empty(R, device="cuda"); f = torch.empty(R, N, 3, device="cuda")
        for _ in range(5):
            eng.energy_force(pos, out_energy=e, out_force=f)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            eng.energy_force(pos, out_energy=e, out_force=f)
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / 100)
        eng.close()
    print(f"N={N} R={R}: node MLPs on tensor cores {out[0]:.4f} ms, on CUDA cores {out[1]:.4f} ms", flush=True)
