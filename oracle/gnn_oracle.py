"""CPU ORACLE for the GNN implicit-solvent energy/force hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product path (gnnimplicitsolvent_b200/) never does.

A flat restatement (plain torch ops on CPU, autograd for the forces, fp32 or fp64) of
    /root/reference/MachineLearning/GNN_Models.py:771-885  (run model composition), :947-1065 (Delta),
    :907-944 (static all-pairs index, eps-distance), :795-804 (r < 0.6 edge selection)
    /root/reference/MachineLearning/GNN_Layers.py:913-1058 (neck tables, Born integral, Born radius),
    :2030-2074, :2121-2140 (GB energy), :2143-2246 (RBF, envelope, edge / node MLP).
Operation order follows the reference so that the fp32 energy is bit-identical to the imported
reference (pinned by tests/test_oracle_vs_reference.py when /root/reference is present, and by the
committed fixtures tests/golden/*.npz otherwise).  The reference ships NO golden vectors or tests
(SURVEY.md §4); parity is pinned against outputs of the reference code itself run on the build box.

Third-party pieces restated from upstream semantics (not vendored in the reference):
  torch_geometric 2.5.1 MessagePassing.propagate  -> gather by edge_index[0] (source j) / [1] (target i),
                                                     index_add onto edge_index[1] in edge order
  torch_cluster 1.6.3 radius_graph                -> strict `<` on squared fp32 distance (data path)
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

OFFSET = 0.0195141
NECK_SCALE = 0.826836
NECK_CUT = 0.68
COULOMB = 138.935485
CUTOFF = 0.6
PROBE = 0.14
DEFAULT_UNIQUE_RADII = [0.14, 0.117, 0.155, 0.15, 0.21, 0.185, 0.18, 0.17, 0.12, 0.13]

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "gnnimplicitsolvent_b200", "data", "neck_tables.npz")


def neck_tables(unique_radii=DEFAULT_UNIQUE_RADII):
    """d0/m0 [U*U] tables: bilinear resampling of the 21x21 GBn2 neck tables at the sorted unique radii
    (GNN_Layers.py:913-972 `createUniqueTable`).  Full tables are data extracted once from the
    reference (tools/extract_neck_tables.py), already rescaled to nm."""
    z = np.load(_DATA)
    d0_full, m0_full = z["d0_nm"].tolist(), z["m0_nm"].tolist()
    radii = list(sorted(set(unique_radii)))
    pos = [(r + OFFSET - 0.1) * 200 for r in radii]
    n = len(radii)
    i1, i2, w1, w2 = [0] * n, [0] * n, [0.0] * n, [0.0] * n
    for i, p in enumerate(pos):
        if p <= 0:
            w1[i] = 1.0
        elif p >= 20:
            i1[i] = 20
            w1[i] = 1.0
        else:
            i1[i] = int(math.floor(p))
            i2[i] = i1[i] + 1
            w1[i] = i2[i] - p
            w2[i] = 1.0 - w1[i]

    def resample(full):
        out = []
        for i in range(n):
            for j in range(n):
                out.append(w1[i] * w1[j] * full[i1[i] * 21 + i1[j]] + w1[i] * w2[j] * full[i1[i] * 21 + i2[j]]
                           + w2[i] * w1[j] * full[i2[i] * 21 + i1[j]] + w2[i] * w2[j] * full[i2[i] * 21 + i2[j]])
        return torch.tensor(out, dtype=torch.float)
    return resample(d0_full), resample(m0_full)


def all_pairs_index(n_atoms: int, n_rep: int) -> torch.Tensor:
    """[2, R*N*(N-1)] static index, laid out [rep][node][con], row0 = node (source j),
    row1 = con (target i)  (GNN_Models.py:919-944 `build_edge_idx`, vectorised)."""
    node = torch.arange(n_atoms).repeat_interleave(n_atoms)
    con = torch.arange(n_atoms).repeat(n_atoms)
    keep = node != con
    node, con = node[keep], con[keep]
    off = (torch.arange(n_rep) * n_atoms).repeat_interleave(node.numel())
    return torch.stack([node.repeat(n_rep) + off, con.repeat(n_rep) + off])


def eps_distance(pos: torch.Tensor, src: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """torch.nn.PairwiseDistance(pos[src], pos[tgt]): ||x_src - x_tgt + 1e-6||_2 (GNN_Models.py:912)."""
    return torch.nn.functional.pairwise_distance(pos.index_select(0, src), pos.index_select(0, tgt))


def neighbor_list_numpy(pos_rnx3: np.ndarray, cutoff: float = CUTOFF, predicate: str = "run"):
    """Bit-exact fp32 neighbour oracle, sorted CSR by target then source.

    predicate "run":  sqrt(fma(dz,dz,fma(dy,dy,dx*dx))) < 0.6f with d = (x_src - x_tgt) + 1e-6f
                      (what torch's CPU PairwiseDistance evaluates -- checked bit-for-bit in
                      tests/test_oracle.py -- followed by GNN_Models.py:795 `edge_attributes < 0.6`)
    predicate "data": sum((x_i - x_j)^2) < r*r   (torch_cluster radius_graph, upstream semantics)
    Returns (row_ptr[R*N+1] int64, col[E] int64 global source index, r[E] float32 eps-distance)."""
    pos = np.ascontiguousarray(pos_rnx3, dtype=np.float32)
    R, N, _ = pos.shape
    d = (pos[:, None, :, :] - pos[:, :, None, :]) + np.float32(1e-6)   # d[r, tgt, src] = x_src - x_tgt + eps
    d64 = d.astype(np.float64)
    acc = (d64[..., 0] * d64[..., 0]).astype(np.float32)
    acc = (d64[..., 1] * d64[..., 1] + acc.astype(np.float64)).astype(np.float32)   # fma rounding
    acc = (d64[..., 2] * d64[..., 2] + acc.astype(np.float64)).astype(np.float32)
    r = np.sqrt(acc)
    if predicate == "run":
        mask = r < np.float32(cutoff)
    else:
        dd = pos[:, :, None, :] - pos[:, None, :, :]
        sq = dd * dd
        d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        mask = d2 < np.float32(cutoff) * np.float32(cutoff)
    mask &= ~np.eye(N, dtype=bool)[None]
    deg = mask.sum(-1).reshape(-1)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    rr, tt, ss = np.nonzero(mask)
    col = (rr * N + ss).astype(np.int64)
    return row_ptr, col, r[rr, tt, ss].astype(np.float32)


class Weights:
    """The 39-entry state dict of the shipped architecture as plain tensors."""

    def __init__(self, state_dict, dtype=torch.float32):
        self.sd = {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach()) for k, v in state_dict.items()}

    def __getitem__(self, k):
        return self.sd[k]


def random_state_dict(seed=0, num_solvents=42, hidden=64, hidden_token=128, unique_radii=DEFAULT_UNIQUE_RADII):
    """Seeded random-init weights of the shipped architecture (torch.nn.Linear / Embedding default init
    scales), with the same 39 keys as MachineLearning/trained_models/GNN.pt["model"]."""
    g = torch.Generator().manual_seed(seed)
    d0, m0 = neck_tables(unique_radii)
    sd = {
        "_num_solvents": torch.tensor(num_solvents, dtype=torch.int32), "_no_batch": torch.tensor(True),
        "aggregate_information._d0": d0, "aggregate_information._m0": m0,
        "calculate_energies._d0": d0.clone(), "calculate_energies._m0": m0.clone(),
        "calculate_energies._soluteDielectric": torch.tensor(1.0), "calculate_energies._solventDielectric": torch.tensor(78.5),
        "lin.weight": torch.zeros(1, 1), "lin.bias": torch.zeros(1),
        "solvent_embedding.weight": torch.randn(num_solvents, hidden, generator=g),
        "gamma_embedding.weight": torch.randn(num_solvents, 1, generator=g),
    }

    def lin(name, o, i):
        k = 1.0 / math.sqrt(i)
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * k
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * k
    for l, (cin, cout) in enumerate([(6, hidden), (2 * hidden, hidden), (2 * hidden, 2)], start=1):
        p = f"interaction{l}"
        sd[p + "._FREQUENCIES"] = math.pi * torch.arange(1, 21, dtype=torch.float)
        lin(p + ".message1", hidden, cin + 20)
        lin(p + ".message2", hidden, hidden)
        lin(p + ".lin1", hidden_token, 2 * hidden)
        lin(p + ".lin2", cout, hidden_token)
    return sd


def _silu(x):
    return torch.nn.functional.silu(x)


def born_radii(x, src, tgt, r, d0, m0, n_unique=10):
    """GBNeck_interaction.forward (GNN_Layers.py:982-1058). x[n,7] = [q, rho, s, alpha, beta, gamma, radindex];
    i = target, j = source.  Returns (B[n], I[n])."""
    dt = x.dtype
    x_i, x_j = x.index_select(0, tgt), x.index_select(0, src)
    or1, or2, sr2 = x_i[:, 1], x_j[:, 1], x_j[:, 2]
    idx1, idx2 = x_i[:, 6].to(torch.int), x_j[:, 6].to(torch.int)
    radius1, radius2 = or1 + OFFSET, or2 + OFFSET
    D = torch.abs(r - sr2)
    L = torch.max(or1, D)
    U = r + sr2
    zero = torch.tensor(0, dtype=dt)
    Ivdw = torch.where((r + sr2 - or1) > 0,
                       0.5 * (1 / L - 1 / U + 0.25 * (r - sr2 ** 2 / r) * (1 / (U ** 2) - 1 / (L ** 2))
                              + 0.5 * torch.log(L / U) / r), zero)
    k = n_unique * idx2 + idx1
    d0k, m0k = d0.index_select(0, k), m0.index_select(0, k)
    Ineck = torch.where((radius1 + radius2 + NECK_CUT - r) > 0,
                        m0k / (1 + 100 * (r - d0k) ** 2 + 0.3 * 1000000 * (r - d0k) ** 6), zero)
    msg = Ivdw + NECK_SCALE * Ineck
    I = torch.zeros(x.size(0), dtype=dt).index_add(0, tgt, msg)
    offset = torch.tensor(OFFSET, dtype=dt)
    radius = x[:, 1] + offset
    psi = I * x[:, 1]
    B = 1 / (1 / x[:, 1] - torch.tanh(x[:, 3] * psi - x[:, 4] * psi ** 2 + x[:, 5] * psi ** 3) / radius)
    return B, I


def rbf(r, freqs, cutoff=CUTOFF):
    """IN_layer_all_swish_2pass.buildsinkernel / envelope (GNN_Layers.py:2194-2208); r [E,1] -> [E,20]."""
    d = r * (1 / cutoff)
    p = 5 + 1
    a, b, c = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
    env = 1.0 / d + a * d ** (p - 1) + b * d ** p + c * d ** (p + 1)
    return env * torch.sin(freqs * d)


def interaction_layer(W, prefix, h, src, tgt, r_e, emb, keep=None):
    """IN_layer_all_swish_2pass_tokens.forward (GNN_Layers.py:2231-2246, message :2186-2192)."""
    e = rbf(r_e, W[prefix + "._FREQUENCIES"])
    z = torch.cat((h.index_select(0, tgt), h.index_select(0, src), e), dim=1)
    u = torch.nn.functional.linear(z, W[prefix + ".message1.weight"], W[prefix + ".message1.bias"])
    w = torch.nn.functional.linear(_silu(u), W[prefix + ".message2.weight"], W[prefix + ".message2.bias"])
    msg = _silu(w)
    a = torch.zeros(h.size(0), msg.size(1), dtype=h.dtype).index_add(0, tgt, msg)
    p = torch.nn.functional.linear(torch.cat((a, emb), dim=1), W[prefix + ".lin1.weight"], W[prefix + ".lin1.bias"])
    o = torch.nn.functional.linear(_silu(p), W[prefix + ".lin2.weight"], W[prefix + ".lin2.bias"])
    if keep is not None:
        keep[prefix] = {"a": a.detach(), "p": p.detach(), "o": o.detach()}
    return o


def gb_energy(Bq, src, tgt, r, diel):
    """GBNeck_energies_no_dielectric.forward (GNN_Layers.py:2123-2140; message :2064-2074, nodewise :2055-2062).
    Bq [n,2] = [B', q]; returns per-atom energies [n,1]."""
    x_i, x_j = Bq.index_select(0, tgt), Bq.index_select(0, src)
    r2 = torch.pow(r, 2)
    B1B2 = x_i[:, 0] * x_j[:, 0]
    qq = x_i[:, 1] * x_j[:, 1]
    f = torch.sqrt(r2 + B1B2 * torch.exp(-r2 / (4 * B1B2)))
    pair = torch.zeros(Bq.size(0), 1, dtype=Bq.dtype).index_add(0, tgt, (qq / f).unsqueeze(1))
    single = (Bq[:, 1] ** 2 / Bq[:, 0]).unsqueeze(1)
    solute = torch.tensor(1, dtype=Bq.dtype)
    return -0.5 * COULOMB * (1 / solute - 1 / diel) * (pair + single)


def evaluate(params, pos, solvent_ids, dielectrics, state_dict, *, dtype=torch.float32, delta=False,
             fraction=0.1, scaling_factor=2.0, forces=True, keep_intermediates=False, predicate="run"):
    """One batched evaluation, the oracle of `gnnis_energy_force`.

    params [N,7] (any float), pos [R,N,3], solvent_ids [R] int, dielectrics [R] float.
    Returns dict: energy_total (scalar), energy_rep [R], e_atom [R*N] (polar), sa_atom [R*N], forces [R,N,3],
    born [R*N], plus csr (row_ptr, col, r) of the GNN graph and optional intermediates."""
    W = Weights(state_dict, dtype)
    params = torch.as_tensor(np.asarray(params), dtype=torch.float32).to(dtype)   # fp32 table first (GNN_Models.py:68-70)
    pos = torch.as_tensor(np.asarray(pos))          # fp32 in the reference; fp64 accepted for finite-difference checks
    R, N, _ = pos.shape
    flat = pos.reshape(R * N, 3).to(dtype).clone().requires_grad_(forces)
    x = params.repeat(R, 1)
    sid = torch.as_tensor(np.asarray(solvent_ids), dtype=torch.int64).repeat_interleave(N)
    diel = torch.as_tensor(np.asarray(dielectrics), dtype=torch.float32).to(dtype).repeat_interleave(N)

    ei = all_pairs_index(N, R)
    src, tgt = ei[0], ei[1]
    r = eps_distance(flat, src, tgt)
    if predicate == "run":      # run models: boolean compaction of the eps-distance (GNN_Models.py:795)
        gnn = r < 0.6
    else:                       # data model: torch_cluster.radius_graph, strict < on the squared fp32 distance
        dd = (flat.detach()[tgt] - flat.detach()[src]).to(torch.float32)
        sq = dd * dd
        gnn = ((sq[:, 0] + sq[:, 1]) + sq[:, 2]) < torch.tensor(0.6, dtype=torch.float32) ** 2
    g_src, g_tgt, g_r = src[gnn], tgt[gnn], r[gnn].unsqueeze(1)

    d0 = W["aggregate_information._d0"]
    m0 = W["aggregate_information._m0"]
    B, I = born_radii(x, src, tgt, r, d0, m0)
    emb = W["solvent_embedding.weight"].index_select(0, sid)
    keep = {} if keep_intermediates else None
    h = torch.stack((B, x[:, 0], x[:, 1]), dim=1)
    h = interaction_layer(W, "interaction1", h, g_src, g_tgt, g_r, emb, keep)
    h = _silu(h)
    h = interaction_layer(W, "interaction2", h, g_src, g_tgt, g_r, emb, keep)
    h = _silu(h)
    h = interaction_layer(W, "interaction3", h, g_src, g_tgt, g_r, emb, keep)
    c_scale, sa_scale = h[:, 0], h[:, 1]
    radius = (x[:, 1] + OFFSET).unsqueeze(1)
    sasa = torch.sigmoid(sa_scale.unsqueeze(1)) * (radius + PROBE) ** 2
    sa = W["gamma_embedding.weight"].index_select(0, sid) * sasa
    Bs = B.unsqueeze(1) * (fraction + torch.sigmoid(c_scale.unsqueeze(1)) * (1 - fraction) * scaling_factor)
    e = gb_energy(torch.cat((Bs, x[:, 0:1]), dim=1), src, tgt, r, diel.unsqueeze(1))
    if delta:
        e = e - gb_energy(torch.stack((B, x[:, 0]), dim=1), src, tgt, r, diel.unsqueeze(1))
    total = e.sum() + sa.sum()
    out = {"energy_total": total.detach(), "e_atom": e.detach().reshape(-1), "sa_atom": sa.detach().reshape(-1),
           "energy_rep": (e + sa).detach().reshape(R, N).sum(1), "born": B.detach(), "born_scaled": Bs.detach().reshape(-1),
           "I": I.detach(), "c_scale": c_scale.detach(), "sa_scale": sa_scale.detach(), "n_edges": int(gnn.sum())}
    if forces:
        (grad,) = torch.autograd.grad(total, flat)
        out["forces"] = (-grad).reshape(R, N, 3)
    if keep_intermediates:
        out["layers"] = keep
    return out


def data_model_graph(pos_flat: np.ndarray, batch: np.ndarray, cutoff: float):
    """Neighbour oracle of the data path (torch_cluster.radius_graph restated; GNN_Models.py:345-357):
    ragged graphs given by `batch`; returns [2,E] (row0 = source, row1 = target) sorted by target, source."""
    x = np.asarray(pos_flat, dtype=np.float32)
    dd = x[:, None, :] - x[None, :, :]
    sq = dd * dd
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    mask = (d2 < np.float32(cutoff) * np.float32(cutoff)) & (batch[:, None] == batch[None, :])
    np.fill_diagonal(mask, False)
    tgt, src = np.nonzero(mask)
    return np.stack([src, tgt])
