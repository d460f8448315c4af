"""Probe: direct launches vs CUDA-graph replay of one evaluation at small sizes (is the graph worth it?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.engine import Engine
z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
for N, R in [(13, 128), (13, 1024), (50, 64), (50, 256), (60, 1024), (50, 4096)]:
    mol = synth.synth_molecule(N, seed=N)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=1)).cuda()
    eng = Engine(mol["params"], sd)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    e = torch.empty(R, device="cuda"); f = torch.empty(R, N, 3, device="cuda")
    for _ in range(5):
        eng.energy_force(pos, out_energy=e, out_force=f)
    torch.cuda.synchronize()
    n = 200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        eng.energy_force(pos, out_energy=e, out_force=f)
    b.record(); torch.cuda.synchronize()
    t_direct = a.elapsed_time(b) / n
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.energy_force(pos, out_energy=e, out_force=f)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.energy_force(pos, out_energy=e, out_force=f)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        g.replay()
    b.record(); torch.cuda.synchronize()
    t_graph = a.elapsed_time(b) / n
    print(f"N={N} R={R}: direct {t_direct:.4f} ms, graph {t_graph:.4f} ms, launches/eval {eng.launch_count() // (n + 7)}", flush=True)
