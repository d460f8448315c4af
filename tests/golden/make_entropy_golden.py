"""Freezes outputs of the reference's own entropy functions (oracle/entropy_ref.py) as golden vectors:
vibrational quasi-RRHO entropies for frequency lists, rotational entropies, and the normal-mode wavenumbers of a
13-atom molecule on the fp64 vacuum-MM oracle potential.  Run in the build container (needs /root/reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnnimplicitsolvent_b200 import synth  # noqa: E402
from oracle import entropy_ref, mm_oracle  # noqa: E402

F = entropy_ref.load_reference_functions()
rng = np.random.default_rng(7)
freq_sets = [np.sort(rng.uniform(20.0, 3500.0, size=n)).astype(np.complex128) for n in (33, 144)]
freq_sets.append(np.array([15.0, 45.0, 99.0, 101.0, 250.0, 1600.0, 3000.0], dtype=np.complex128))
sv = [F["calculate_entropy_from_frequencies"](f) for f in freq_sets]

mol = synth.synth_molecule(13, seed=9)
ff = synth.synth_forcefield(mol)
force_fn = lambda p: mm_oracle.evaluate(ff, p[None])["forces"][0].numpy()
# relax first (plain gradient descent is enough for a golden geometry near a minimum)
import scipy.optimize  # noqa: E402
fun = lambda v: (float(mm_oracle.evaluate(ff, v.reshape(1, 13, 3), forces=False)["energy_rep"][0]),
                 -force_fn(v.reshape(13, 3)).ravel())
res = scipy.optimize.minimize(fun, mol["pos"].ravel(), jac=True, method="BFGS", options={"gtol": 1e-6, "maxiter": 5000})
xmin = res.x.reshape(13, 3)
sim = entropy_ref.FakeSimulation(xmin, mol["masses"], force_fn)
hess = F["build_mass_weighted_hessian"](sim, delta=0.0001)
freqs, _ = F["normal_modes"](sim, delta=0.0001)
srot = F["calculate_rotational_entropy"](sim)
np.savez(os.path.join(ROOT, "tests", "golden", "entropy_golden.npz"),
         freq0=freq_sets[0], freq1=freq_sets[1], freq2=freq_sets[2], sv=np.array(sv, dtype=np.float64),
         xmin=xmin, masses=np.asarray(mol["masses"], dtype=np.float64), hessian=hess, wavenumbers=freqs,
         srot=float(srot), svib=float(F["calculate_entropy_from_frequencies"](freqs)))
print("sv", sv, "srot", float(srot), "lowest modes", freqs[:4], "highest", freqs[-2:])
