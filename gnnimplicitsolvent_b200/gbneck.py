"""GB-Neck2 neck tables: bilinear resampling of the 21x21 (d0, m0) tables at the model's sorted unique radii.

Mirrors what the reference layers register as buffers `_d0/_m0`
(MachineLearning/GNN_Layers.py:913-972 `createUniqueTable`); the kernels index them with
k = n_unique * radindex_source + radindex_target (GNN_Layers.py:974-980)."""
from __future__ import annotations

import os

import numpy as np

OFFSET = 0.0195141
DEFAULT_UNIQUE_RADII = [0.14, 0.117, 0.155, 0.15, 0.21, 0.185, 0.18, 0.17, 0.12, 0.13]
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "neck_tables.npz")


def neck_tables(unique_radii=DEFAULT_UNIQUE_RADII):
    """Returns (d0, m0, n_unique): float32 arrays of n_unique**2 entries (nm units)."""
    z = np.load(_DATA)
    full = {"d0": z["d0_nm"], "m0": z["m0_nm"]}
    radii = sorted(set(float(r) for r in unique_radii))
    n = len(radii)
    lo = np.zeros(n, dtype=np.int64)
    hi = np.zeros(n, dtype=np.int64)
    wlo = np.zeros(n)
    whi = np.zeros(n)
    for i, r in enumerate(radii):
        p = (r + OFFSET - 0.1) * 200
        if p <= 0:
            wlo[i] = 1.0
        elif p >= 20:
            lo[i] = 20
            wlo[i] = 1.0
        else:
            lo[i] = int(np.floor(p))
            hi[i] = lo[i] + 1
            wlo[i] = hi[i] - p
            whi[i] = 1.0 - wlo[i]
    out = {}
    for key, tab in full.items():
        t = tab.reshape(21, 21)
        res = np.zeros((n, n))
        for i in range(n):
            for j in range(n):
                res[i, j] = (wlo[i] * wlo[j] * t[lo[i], lo[j]] + wlo[i] * whi[j] * t[lo[i], hi[j]]
                             + whi[i] * wlo[j] * t[hi[i], lo[j]] + whi[i] * whi[j] * t[hi[i], hi[j]])
        out[key] = res.reshape(-1).astype(np.float32)
    return out["d0"], out["m0"], n


# GBn2 per-element standard parameters: (radius nm, screen, alpha, beta, gamma).  Restated from OpenMM 8.0.0
# `openmm.app.internal.customgbforces` (GBSAGBn2Force.getStandardParameters: mbondi3 radii for small molecules =
# mbondi2, hydrogens bonded to nitrogen 0.13 nm), which the reference calls at MachineLearning/GNN_Trainer.py:854-857.
# UNPINNED: OpenMM is absent here; verify where it is installed before relying on elements beyond H, C, N, O, S.
GBN2_ELEMENT = {
    "H": (0.12, 1.425952, 0.788440, 0.798699, 0.437334),
    "C": (0.17, 1.058554, 0.733756, 0.506378, 0.205844),
    "N": (0.155, 0.733599, 0.503364, 0.316828, 0.192915),
    "O": (0.15, 1.061039, 0.867814, 0.876635, 0.387882),
    "F": (0.15, 0.5, 1.0, 0.8, 4.851),
    "Si": (0.21, 0.5, 1.0, 0.8, 4.851),
    "P": (0.185, 0.5, 1.0, 0.8, 4.851),
    "S": (0.18, -0.703469, 0.867814, 0.876635, 0.387882),
    "Cl": (0.17, 0.5, 1.0, 0.8, 4.851),
}


def gbneck2_parameters(elements, bonds, charges, unique_radii=DEFAULT_UNIQUE_RADII):
    """The [N,7] parameter table of the run models, `[q, rho, s, alpha, beta, gamma, radindex]`, from element symbols,
    the bond list (for the H-on-N radius) and partial charges -- what Trainer.get_gbneck2_param_small_molecules
    (MachineLearning/GNN_Trainer.py:829-869) extracts from an OpenMM system:
        rho = radius - 0.0195141 (:865-866), s = screen * rho (:867),
        radindex = position of the RAW radius in `unique_radii` in the order GIVEN (:861-864) -- for the default list
        that is the unsorted order of GNN_Models.py:38, the quirk the shipped weights were used with.
    Returns (params float64 [N,7], unique_radii)."""
    n = len(elements)
    nbr_n = [False] * n
    for i, j in bonds:
        if elements[j] == "N":
            nbr_n[i] = True
        if elements[i] == "N":
            nbr_n[j] = True
    out = np.zeros((n, 7), dtype=np.float64)
    lookup = {float("%.4f" % r): k for k, r in enumerate(unique_radii)}
    for i, el in enumerate(elements):
        radius, screen, a, b, g = GBN2_ELEMENT[el]
        if el == "H" and nbr_n[i]:
            radius = 0.13
        key = float("%.4f" % radius)
        if key not in lookup:
            raise KeyError(f"radius {radius} of element {el} is not in unique_radii")
        rho = radius - OFFSET
        out[i] = [charges[i], rho, screen * rho, a, b, g, lookup[key]]
    return out, list(unique_radii)
