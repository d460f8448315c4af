"""Generates tests/golden/*.npz by running the UNMODIFIED reference model
(/root/reference/MachineLearning/GNN_Models.py: GNN3_Multisolvent_embedding_run_multiple :700,
 ..._Delta :947, ..._split :1068) on CPU through oracle/ref_harness.py, with the shipped checkpoint
MachineLearning/trained_models/GNN.pt.  Run on the build box only (the reference tree does not travel):
    python tests/golden/make_golden.py [case names]
Outputs (committed):
    gnn_weights.npz          the 39-entry state dict of GNN.pt["model"] (weights are data, not source)
    case_<name>.npz          inputs (params, pos, solvent ids, dielectrics) + reference outputs:
                             energy_total (run), energy_total_delta (Delta), e_atom / sa_atom (split),
                             forces (autograd on the run model), forces_delta, Born radii,
                             GNN edge list (the reference's own `edge_attributes < 0.6` selection).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402
from gnnimplicitsolvent_b200 import synth  # noqa: E402

CASES = {
    # name: (N, R, solvent ids, jitter sigma)
    "c1_n13_dmso": (13, 8, [3] * 8, 0.01),
    "mixed_n21": (21, 5, [0, 3, 7, 11, 38], 0.01),
    "c2_n50_water": (50, 4, [0] * 4, 0.01),
    "mixed_n50": (50, 6, [1, 5, 9, 20, 30, 38], 0.02),
    "n100_mixed": (100, 2, [0, 12], 0.01),
    # round 2: two sulfur atoms (negative GBn2 screening scale: s_j < 0) -- molecule seed 2 instead of N
    "sulfur_n30": (30, 4, [0, 3, 9, 22], 0.01),
    # round 2: in every conformer one pair sits inside the 1e-6-wide eps band of the 0.6 nm cutoff, so that exactly ONE
    # of its two directed edges exists (r(j->i) < 0.6 <= r(i->j)); see eps_band_conformers
    "epsband_n21": (21, 4, [0, 3, 7, 38], 0.01),
}
MOL_SEED = {"sulfur_n30": 2}


def eps_band_conformers(pos, cutoff=0.6):
    """Moves one atom of every conformer along a pair axis until the fp32 eps-distances of the two directions
    (|x_src - x_tgt + 1e-6|, GNN_Models.py:912) straddle the cutoff: the edge list becomes asymmetric there."""
    from oracle import gnn_oracle as O
    pos = pos.copy()
    R, N, _ = pos.shape
    for r in range(R):
        d = np.linalg.norm(pos[r][:, None, :] - pos[r][None, :, :], axis=2)
        d[np.tril_indices(N)] = 99.0
        i, j = np.unravel_index(np.argmin(np.abs(d - cutoff)), d.shape)
        axis = (pos[r, j] - pos[r, i]).astype(np.float64)
        axis /= np.linalg.norm(axis)
        found = False
        for k in range(-4000, 4000):           # steps of 2.5e-8 nm around the cutoff
            trial = pos[r].copy()
            trial[j] = (pos[r, i].astype(np.float64) + axis * (cutoff + k * 2.5e-8)).astype(np.float32)
            rp, col, _ = O.neighbor_list_numpy(trial[None])
            has_ji = j in col[rp[i]:rp[i + 1]]     # edge j -> i (row of target i)
            has_ij = i in col[rp[j]:rp[j + 1]]
            if has_ji != has_ij:
                pos[r] = trial
                found = True
                break
        assert found, "no asymmetric placement found"
    return pos


def main():
    torch.set_num_threads(1)
    sd = torch.load(H.checkpoint_path(), weights_only=True, map_location="cpu")["model"]
    if not sys.argv[1:]:
        np.savez_compressed(os.path.join(HERE, "gnn_weights.npz"), **{k: v.numpy() for k, v in sd.items()})
    only = sys.argv[1:]
    for name, (N, R, sids, sigma) in CASES.items():
        if only and name not in only:
            continue
        mol = synth.synth_molecule(N, seed=MOL_SEED.get(name, N))
        pos = synth.conformers(mol["pos"], R, sigma, seed=1000 * N + R)
        if name.startswith("epsband"):
            pos = eps_band_conformers(pos)
        diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
        out = {"params": mol["params"], "pos": pos, "solvent_ids": np.array(sids, dtype=np.int64),
               "dielectrics": np.array(diel, dtype=np.float32)}
        minus1, d785 = torch.tensor(-1), torch.tensor(78.5)

        run = H.build_reference_run_model(mol["params"].tolist(), sd, R, sids, diel)
        p = torch.tensor(pos.reshape(-1, 3), requires_grad=True)
        e = run(p, minus1, d785)
        e.backward()
        out["energy_total"] = e.detach().numpy()
        out["forces"] = (-p.grad).reshape(R, N, 3).numpy()
        # the reference's own edge selection (GNN_Models.py:794-804)
        with torch.no_grad():
            _, ei, ea = run.build_graph(p.detach())
            keep = (ea < 0.6).squeeze()
            src, tgt, r = ei[0][keep].numpy(), ei[1][keep].numpy(), ea[keep].numpy().reshape(-1)
            order = np.lexsort((src, tgt))
            out["edge_src"], out["edge_tgt"], out["edge_r"] = src[order], tgt[order], r[order]
            Bc = run.aggregate_information(x=run._batch_gbparameters, edge_index=ei, edge_attributes=ea)
            out["born"] = Bc[:, 0].numpy()

        split = H.build_reference_run_model(mol["params"].tolist(), sd, R, sids, diel,
                                            cls="GNN3_Multisolvent_embedding_run_multiple_split")
        with torch.no_grad():
            ea_, sa_ = split(torch.tensor(pos.reshape(-1, 3)), minus1, d785)
        out["e_atom"], out["sa_atom"] = ea_.numpy().reshape(-1), sa_.numpy().reshape(-1)

        delta = H.build_reference_run_model(mol["params"].tolist(), sd, R, sids, diel,
                                            cls="GNN3_Multisolvent_embedding_run_multiple_Delta")
        p2 = torch.tensor(pos.reshape(-1, 3), requires_grad=True)
        e2 = delta(p2)
        e2.backward()
        out["energy_total_delta"] = e2.detach().numpy()
        out["forces_delta"] = (-p2.grad).reshape(R, N, 3).numpy()

        # single-solvent override path (solvent_model != -1, GNN_Models.py:780-786) on replica 0 alone
        one = H.build_reference_run_model(mol["params"].tolist(), sd, 1, [0], [78.5])
        p3 = torch.tensor(pos[0], requires_grad=True)
        e3 = one(p3, torch.tensor(sids[0]), torch.tensor(diel[0]))
        e3.backward()
        out["energy_rep0_override"] = e3.detach().numpy()
        out["forces_rep0_override"] = (-p3.grad).numpy()

        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
        rep = synth.geometry_report(pos)
        print(name, "E", float(e), "Edelta", float(e2), "edges", len(src), rep)


if __name__ == "__main__":
    main()
