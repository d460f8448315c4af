"""Development tool (GPU box): time of the neighbour stage (Born radii + graph build) with the all-pairs and the cell-list
back-end over molecule sizes and shapes, plus a byte comparison of the CSR arrays.  One JSON line per point.

usage: python tools/neighbor_backends.py > gpurun_out/neighbor_backends.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gnnimplicitsolvent_b200 import _lib, synth  # noqa: E402
from gnnimplicitsolvent_b200.engine import Engine  # noqa: E402


def chain(N, step, seed):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(N, 3))
    d[:, 0] += 1.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (np.cumsum(d * step, axis=0) + rng.normal(scale=0.05, size=(N, 3))).astype(np.float32)


def globule(N, seed):
    """Atoms at liquid-like density (~100 atoms / nm^3) in a ball: what a folded chain of N atoms looks like."""
    rng = np.random.default_rng(seed)
    rad = (3.0 * N / (4.0 * np.pi * 100.0)) ** (1.0 / 3.0)
    x = rng.normal(size=(4 * N, 3))
    x *= (rng.random(4 * N) ** (1.0 / 3.0) * rad / np.linalg.norm(x, axis=1))[:, None]
    return x[:N].astype(np.float32)


z = np.load(os.path.join(ROOT, "tests", "golden", "gnn_weights.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N, R, shape in [(50, 4096, "synth"), (100, 1024, "synth"), (200, 512, "synth"), (200, 512, "chain"), (512, 128, "globule"),
                    (512, 128, "chain"), (1024, 32, "globule"), (1024, 32, "chain")]:
    mol = synth.synth_molecule(min(N, 200), seed=N)
    if shape != "synth":   # parameter rows repeated: building a self-avoiding 1024-atom molecule takes minutes on the host
        mol = {"params": np.resize(mol["params"], (N, 7))}
    if shape == "synth":
        pos_h = synth.conformers(mol["pos"], R, 0.01, seed=1)
    elif shape == "chain":
        pos_h = np.stack([chain(N, 0.12, 7 * N + r) for r in range(R)])
    else:
        pos_h = np.stack([globule(N, 7 * N + r) for r in range(R)])
    pos = torch.from_numpy(pos_h).cuda()
    eng = Engine(mol["params"], sd)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    a = eng.build_neighbors(pos, backend="allpairs")
    b = eng.build_neighbors(pos, backend="cell")
    same = all(torch.equal(x.view(torch.int32), y.view(torch.int32)) for x, y in zip(a, b))
    out = {"N": N, "R": R, "shape": shape, "edges_per_atom": round(a[1].numel() / (N * R), 2), "identical": bool(same)}
    for name, fl in (("allpairs", _lib.FLAG_ALL_PAIRS), ("cell", _lib.FLAG_CELL_LIST)):
        acc, tot = 0.0, 0.0
        for it in range(8):
            flush.fill_(1)
            st = eng.profile_eval(pos, extra_flags=fl)
            if it >= 2:
                acc += st["neighbor_born"]
                tot += sum(st.values())
        out[name + "_neighbor_ms"] = round(acc / 6, 4)
        out[name + "_eval_ms"] = round(tot / 6, 4)
    out["path"] = eng.describe()
    print(json.dumps(out), flush=True)
    eng.close()
