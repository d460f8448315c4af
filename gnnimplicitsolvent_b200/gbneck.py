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
