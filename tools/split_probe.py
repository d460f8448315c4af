import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
z = np.load("/root/repo/tests/golden/gnn_weights.npz")
sd = {k: torch.from_numpy(z[k]) for k in z.files}
for N, R in [(100, 1), (100, 1024), (128, 1024), (70, 2048), (50, 8)]:
    mol = synth.synth_molecule(N, seed=N)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=1)).cuda()
    res = {}
    for sp in ("1", "2", "16"):
        os.environ["GNNIS_SPLIT"] = sp
        eng = Engine(mol["params"], sd)
        eng.set_replicas(R, [r % 39 for r in range(R)], [78.5] * R)
        e = torch.empty(R, device="cuda"); f = torch.empty(R, N, 3, device="cuda")
        for _ in range(5):
            eng.energy_force(pos, out_energy=e, out_force=f)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            eng.energy_force(pos, out_energy=e, out_force=f)
        b.record(); torch.cuda.synchronize()
        res[sp] = (a.elapsed_time(b) / 50, e.clone(), f.clone())
        eng.close()
    de = float((res["1"][1] - res["16"][1]).abs().max() / res["1"][1].abs().max())
    df = float(((res["1"][2] - res["16"][2]) ** 2).sum().sqrt() / (res["1"][2] ** 2).sum().sqrt())
    print(f"N={N} R={R}: split off {res['1'][0]:.4f} ms, max 2 {res['2'][0]:.4f} ms, max 16 {res['16'][0]:.4f} ms, E diff {de:.2e}, F diff {df:.2e}", flush=True)
