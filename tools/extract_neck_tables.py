"""One-off: pull the 21x21 GBn2 neck tables (d0, m0; already rescaled to nm) out of the reference layer
(GNN_Layers.py:1075-1963, attribute `_GBNeck_energies__d0/__m0`) into gnnimplicitsolvent_b200/data/neck_tables.npz.
Runs only where /root/reference exists. The tables are published GBn2 force-field DATA (Nguyen et al. 2013)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

M = H.import_reference_models()
layer = M.GBNeck_energies([[0, 0.15, 0, 0, 0, 0, 0]], "cpu", unique_radii=M.DEFAULT_UNIQUE_RADII)
np.savez(os.path.join(ROOT, "gnnimplicitsolvent_b200", "data", "neck_tables.npz"),
         d0_nm=np.array(layer._GBNeck_energies__d0, dtype=np.float64),
         m0_nm=np.array(layer._GBNeck_energies__m0, dtype=np.float64))
print("written")
