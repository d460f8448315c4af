"""Synthetic molecules / conformer batches; numpy only, deterministic per seed.

RDKit, OpenFF and OpenMM are absent here and on the GPU box, so the benchmark and the parity tests use
a self-avoiding branched heavy-atom walk decorated with hydrogens, with GBn2 per-element parameters
laid out exactly as the reference's parameter table:
    [charge, rho = radius - 0.0195141, s = scale * rho, alpha, beta, gamma, radindex]
(reference: MachineLearning/GNN_Trainer.py:829-869; radindex indexes the UNSORTED
DEFAULT_UNIQUE_RADII list, MachineLearning/GNN_Models.py:38 -- the quirk is kept on purpose).
"""
from __future__ import annotations

import numpy as np

OFFSET = 0.0195141
DEFAULT_UNIQUE_RADII = [0.14, 0.117, 0.155, 0.15, 0.21, 0.185, 0.18, 0.17, 0.12, 0.13]

# element -> (radius nm, scale, alpha, beta, gamma)   GBn2 table (values only need to be identical on both sides)
GBN2 = {
    "H": (0.12, 1.425952, 0.788440, 0.798699, 0.437334),
    "HN": (0.13, 1.425952, 0.788440, 0.798699, 0.437334),
    "C": (0.17, 1.058554, 0.733756, 0.506378, 0.205844),
    "N": (0.155, 0.733599, 0.503364, 0.316828, 0.192915),
    "O": (0.15, 1.061039, 0.867814, 0.876635, 0.387882),
    "S": (0.18, -0.703469, 0.867814, 0.876635, 0.387882),
}
MASS = {"H": 1.008, "HN": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "S": 32.06}
_VALENCE = {"C": 4, "N": 3, "O": 2, "S": 2}


def _rand_unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def _try_place(rng, pos, parent, nbrs_of_parent, bond, min_nb, min_angle_deg, s_mask, tries=200):
    """Proposes a position bonded to `parent`: bond length `bond`, angle to the parent's other bonds
    >= min_angle_deg, distance to all non-bonded atoms >= min_nb (>= 0.15 nm to any sulfur)."""
    p0 = pos[parent]
    cos_max = np.cos(np.deg2rad(min_angle_deg))
    P = np.asarray(pos)
    for _ in range(tries):
        d = _rand_unit(rng)
        ok = True
        for nb in nbrs_of_parent:
            u = pos[nb] - p0
            u = u / np.linalg.norm(u)
            if float(np.dot(u, d)) > cos_max:
                ok = False
                break
        if not ok:
            continue
        cand = p0 + bond * d
        dist = np.linalg.norm(P - cand, axis=1)
        dist[parent] = np.inf
        lim = np.full(len(P), min_nb)
        # 1-3 neighbours (bonded to the parent) may be closer than the non-bonded limit
        for nb in nbrs_of_parent:
            lim[nb] = 0.20 if bond > 0.12 else 0.16
        if np.all(dist >= lim):
            return cand
    return None


def synth_molecule(n_atoms: int, seed: int = 0):
    """Returns dict(params[N,7] float64, pos[N,3] float64 nm, elements list, masses[N], bonds list).

    ~50 % hydrogens; heavy elements C .60 / N .15 / O .20 / S .05; charges ~ N(0, 0.3^2) with zero net
    charge. Every distance involving sulfur is >= 0.15 nm (negative GBn2 scale => U = r + s_j must stay
    well above 0)."""
    rng = np.random.default_rng(seed)
    n_h = n_atoms // 2
    n_heavy = n_atoms - n_h
    for _attempt in range(50):
        elements, pos, bonds = [], [], []
        nbrs = {}
        e0 = "C"
        elements.append(e0)
        pos.append(np.zeros(3))
        nbrs[0] = []
        fail = False
        while len(pos) < n_heavy:
            cands = [i for i in range(len(pos)) if len(nbrs[i]) < _VALENCE[elements[i]] - (1 if elements[i] == "C" else 0)]
            if not cands:
                cands = [i for i in range(len(pos)) if len(nbrs[i]) < _VALENCE[elements[i]]]
            if not cands:
                fail = True
                break
            # prefer recently added atoms => chain-like with branches
            w = np.array([1.0 + 3.0 * (i / max(1, len(pos) - 1)) for i in cands])
            parent = int(rng.choice(cands, p=w / w.sum()))
            el = str(rng.choice(["C", "N", "O", "S"], p=[0.60, 0.15, 0.20, 0.05]))
            bond = 0.18 if "S" in (el, elements[parent]) else 0.15
            c = _try_place(rng, pos, parent, nbrs[parent], bond, 0.25, 95.0, None)
            if c is None:
                continue
            idx = len(pos)
            pos.append(c)
            elements.append(el)
            nbrs[idx] = [parent]
            nbrs[parent].append(idx)
            bonds.append((parent, idx))
        if fail:
            continue
        # hydrogens on free valences, round-robin over heavy atoms
        placed = 0
        guard = 0
        order = list(range(n_heavy))
        while placed < n_h and guard < 20 * n_atoms:
            guard += 1
            free = [i for i in order if len(nbrs[i]) < _VALENCE[elements[i]] and elements[i] != "S"]
            if not free:
                free = [i for i in order if elements[i] == "C"]  # over-saturate a carbon rather than fail
                if not free:
                    free = order
            parent = int(rng.choice(free))
            c = _try_place(rng, pos, parent, nbrs[parent], 0.109, 0.19, 90.0, None)
            if c is None:
                continue
            idx = len(pos)
            pos.append(c)
            elements.append("HN" if elements[parent] == "N" else "H")
            nbrs[idx] = [parent]
            nbrs[parent].append(idx)
            bonds.append((parent, idx))
            placed += 1
        if placed == n_h:
            break
    else:
        raise RuntimeError(f"synth_molecule({n_atoms}, {seed}) failed")

    pos = np.asarray(pos, dtype=np.float64)
    q = rng.normal(0.0, 0.3, size=n_atoms)
    q -= q.mean()
    params = np.zeros((n_atoms, 7), dtype=np.float64)
    for i, el in enumerate(elements):
        radius, scale, a, b, g = GBN2[el]
        rho = radius - OFFSET
        params[i] = [q[i], rho, scale * rho, a, b, g, DEFAULT_UNIQUE_RADII.index(radius)]
    masses = np.array([MASS[e] for e in elements])
    return {"params": params, "pos": pos, "elements": elements, "masses": masses, "bonds": bonds}


def conformers(base_pos: np.ndarray, n_rep: int, sigma: float = 0.01, seed: int = 0) -> np.ndarray:
    """R copies = base geometry + per-atom Gaussian jitter (nm). Returns float32 [R, N, 3]."""
    rng = np.random.default_rng(seed)
    x = base_pos[None, :, :] + rng.normal(0.0, sigma, size=(n_rep,) + base_pos.shape)
    return x.astype(np.float32)


def geometry_report(pos: np.ndarray, cutoff: float = 0.6) -> dict:
    """min pair distance and min |r - cutoff| margin over all intra-replica pairs (pos [R,N,3] fp32)."""
    pos = np.asarray(pos, dtype=np.float32)
    if pos.ndim == 2:
        pos = pos[None]
    d = pos[:, :, None, :] - pos[:, None, :, :] + np.float32(1e-6)
    r = np.sqrt((d * d).sum(-1, dtype=np.float32))
    n = pos.shape[1]
    off = ~np.eye(n, dtype=bool)
    rr = r[:, off]
    return {"min_dist": float(rr.min()), "min_cutoff_margin": float(np.abs(rr - np.float32(cutoff)).min()),
            "edges_per_replica": float((rr < np.float32(cutoff)).sum() / pos.shape[0])}


SOLVENT_DIELECTRICS = [
    78.5, 4.81, 33.0, 47.24, 36.12, 4.27, 25.3, 38.25, 8.93, 2.38, 2.28, 1.88, 36.64, 21.01, 6.20, 2.22,
    35.6, 29.6, 4.5, 20.18, 2.03, 13.26, 7.52, 6.20, 43.3, 37.27, 6.10, 32.55, 10.3, 2.024, 46.53, 2.24,
    7.30, 26.74, 9.22, 2.104, 29.7, 25.9, 2.56,
]  # Simulation/solvents.yml:2-40, indexed by solvent_id 0..38


# GAFF-like Lennard-Jones parameters (sigma nm, epsilon kJ/mol) per element of the synthetic molecules
LJ = {"H": (0.2650, 0.0657), "HN": (0.2650, 0.0657), "C": (0.3400, 0.4577), "N": (0.3250, 0.7113),
      "O": (0.2960, 0.8786), "S": (0.3564, 1.0460)}


def synth_forcefield(mol, coulomb14=1.0 / 1.2, lj14=0.5):
    """A complete vacuum force field for a `synth_molecule`, in the layout of an OpenMM System built by OpenFF
    (HarmonicBondForce, HarmonicAngleForce, PeriodicTorsionForce, NonbondedForce with 1-2 / 1-3 exclusions and scaled
    1-4 exceptions; reference: Simulation/Simulator.py:2036-2143 for how the non-bonded part is re-expressed).
    Equilibrium values are the base geometry's; force constants are typical small-molecule values."""
    pos, elements, bonds = mol["pos"], mol["elements"], [tuple(b) for b in mol["bonds"]]
    n = len(elements)
    nbrs = {i: [] for i in range(n)}
    for i, j in bonds:
        nbrs[i].append(j)
        nbrs[j].append(i)
    dist = lambda a, b: float(np.linalg.norm(pos[a] - pos[b]))
    b_idx = np.array(bonds, dtype=np.int32).reshape(-1, 2)
    b_r0 = np.array([dist(i, j) for i, j in bonds], dtype=np.float32)
    b_k = np.array([3.0e5 if elements[i].startswith("H") or elements[j].startswith("H") else 2.5e5 for i, j in bonds],
                   dtype=np.float32)
    a_idx, a_t0 = [], []
    for j in range(n):
        nb = sorted(nbrs[j])
        for x in range(len(nb)):
            for y in range(x + 1, len(nb)):
                i, k = nb[x], nb[y]
                u, v = pos[i] - pos[j], pos[k] - pos[j]
                c = float(np.dot(u, v) / (np.linalg.norm(u) * np.linalg.norm(v)))
                a_idx.append((i, j, k))
                # a harmonic angle is singular at 180 deg unless theta0 = 180 deg: keep equilibrium angles well away
                a_t0.append(min(np.arccos(np.clip(c, -1.0, 1.0)), np.deg2rad(150.0)))
    t_idx, t_n, t_ph, t_k = [], [], [], []
    for (j, k) in bonds:
        for i in sorted(nbrs[j]):
            if i == k:
                continue
            for l in sorted(nbrs[k]):
                if l == j or l == i:
                    continue
                t_idx.append((i, j, k, l))
                if len(t_idx) % 3 == 0:
                    t_n.append(2), t_ph.append(np.pi), t_k.append(2.0)
                else:
                    t_n.append(3), t_ph.append(0.0), t_k.append(0.6)
    # topological distances for the exception list
    sep = np.full((n, n), 99, dtype=np.int64)
    for s in range(n):
        sep[s, s] = 0
        frontier = [s]
        d = 0
        while frontier and d < 3:
            d += 1
            nxt = []
            for u in frontier:
                for v in nbrs[u]:
                    if sep[s, v] > d:
                        sep[s, v] = d
                        nxt.append(v)
            frontier = nxt
    q = np.asarray(mol["params"][:, 0], dtype=np.float64)
    sig = np.array([LJ[e][0] for e in elements])
    eps = np.array([LJ[e][1] for e in elements])
    x_idx, x_qq, x_s, x_e = [], [], [], []
    for i in range(n):
        for j in range(i + 1, n):
            if sep[i, j] in (1, 2):
                x_idx.append((i, j)), x_qq.append(0.0), x_s.append(0.5 * (sig[i] + sig[j])), x_e.append(0.0)
            elif sep[i, j] == 3:
                x_idx.append((i, j)), x_qq.append(coulomb14 * q[i] * q[j])
                x_s.append(0.5 * (sig[i] + sig[j])), x_e.append(lj14 * np.sqrt(eps[i] * eps[j]))
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    return {
        "bonds": {"idx": b_idx, "r0": b_r0, "k": b_k},
        "angles": {"idx": np.array(a_idx, dtype=np.int32).reshape(-1, 3), "theta0": f32(a_t0), "k": f32([400.0] * len(a_idx))},
        "torsions": {"idx": np.array(t_idx, dtype=np.int32).reshape(-1, 4), "periodicity": np.array(t_n, dtype=np.int32),
                     "phase": f32(t_ph), "k": f32(t_k)},
        "charge": f32(q), "sigma": f32(sig), "epsilon": f32(eps),
        "exceptions": {"idx": np.array(x_idx, dtype=np.int32).reshape(-1, 2), "chargeprod": f32(x_qq), "sigma": f32(x_s),
                       "epsilon": f32(x_e)},
    }


def synth_mol_object(n_atoms: int, n_conformers: int, seed: int = 0, sigma: float = 0.01, with_forcefield: bool = True):
    """A `Simulator.Molecule` (the stand-in for an RDKit molecule with embedded conformers) of a synthetic molecule:
    conformers = jittered copies of the base geometry, parameters and vacuum force field as above."""
    from .Simulator import Molecule  # noqa: PLC0415

    mol = synth_molecule(n_atoms, seed=seed)
    conf = conformers(mol["pos"], n_conformers, sigma, seed=seed + 1)
    return Molecule(mol["elements"], conf, mol["params"], synth_forcefield(mol) if with_forcefield else None,
                    masses=mol["masses"], smiles=f"SYNTH{n_atoms}")
