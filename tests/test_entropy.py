"""Entropy path against golden vectors produced by the reference's OWN functions
(Simulation/helper_functions.py:1551-1885, run through oracle/entropy_ref.py by tests/golden/make_entropy_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "entropy_golden.npz")))


def test_vibrational_entropy_matches_reference(gold):
    from gnnimplicitsolvent_b200 import entropy
    for k in range(3):
        s = entropy.vibrational_entropy(gold[f"freq{k}"])
        assert abs(s - gold["sv"][k]) < 1e-9 * abs(gold["sv"][k])
    assert abs(entropy.vibrational_entropy(gold["wavenumbers"]) - float(gold["svib"])) < 1e-9


def test_rotational_entropy_matches_reference(gold):
    from gnnimplicitsolvent_b200 import entropy
    s = entropy.rotational_entropy(gold["xmin"], gold["masses"])
    assert abs(s - float(gold["srot"])) < 1e-9 * float(gold["srot"])


def test_normal_modes_from_reference_hessian(gold):
    from gnnimplicitsolvent_b200 import entropy
    nu = entropy.normal_mode_wavenumbers(torch.from_numpy(gold["hessian"])[None], 13)[0]
    np.testing.assert_allclose(nu.real, gold["wavenumbers"].real, rtol=1e-9)
    assert np.all(nu.imag == 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists on the build box")
def test_reference_functions_still_reproduce_golden(gold):
    from oracle import entropy_ref
    F = entropy_ref.load_reference_functions()
    assert abs(float(F["calculate_entropy_from_frequencies"](gold["freq2"])) - gold["sv"][2]) < 1e-12


@pytest.mark.gpu
def test_batched_hessian_vs_reference(gold):
    """All 6N displaced copies in ONE batched evaluation reproduce the reference's coordinate-by-coordinate Hessian
    (fp32 forces here, fp64 oracle forces there) and its wavenumbers / entropy."""
    from gnnimplicitsolvent_b200 import entropy, synth
    from gnnimplicitsolvent_b200.engine import Engine
    mol = synth.synth_molecule(13, seed=9)
    eng = Engine(mol["params"], None)
    eng.set_forcefield(synth.synth_forcefield(mol))
    hess = entropy.mass_weighted_hessians(eng, gold["xmin"][None], gold["masses"], vacuum_only=True)
    ref = gold["hessian"]
    err = float(np.abs(hess[0].cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 2e-4, err
    nu = entropy.normal_mode_wavenumbers(hess, 13)[0]
    assert np.all(nu.imag == 0)
    np.testing.assert_allclose(nu.real[3:], gold["wavenumbers"].real[3:], rtol=2e-3)
    np.testing.assert_allclose(nu.real[:3], gold["wavenumbers"].real[:3], rtol=3e-2)
    s, _ = entropy.conformer_entropies(eng, gold["xmin"][None], gold["masses"], vacuum_only=True)
    assert abs(s[0] - (float(gold["svib"]) + float(gold["srot"]))) < 0.1     # kJ/mol


@pytest.mark.gpu
def test_calculate_entropy_full_hamiltonian(weights):
    """Minimise + entropies for several conformers on the GNN + vacuum Hamiltonian: finite, bounded, reproducible."""
    from gnnimplicitsolvent_b200 import entropy, synth
    from gnnimplicitsolvent_b200.engine import Engine
    mol = synth.synth_molecule(21, seed=3)
    eng = Engine(mol["params"], weights)
    eng.set_forcefield(synth.synth_forcefield(mol))
    pos = synth.conformers(mol["pos"], 6, 0.01, seed=4)
    s, g = entropy.calculate_entropy(eng, pos, mol["masses"], solvent_id=3, dielectric=47.24, grad_tol=5.0)
    assert s.shape == (6,) and np.all(np.isfinite(s)) and np.all(np.isfinite(g))
    assert np.all(s > 20.0) and np.all(s < 400.0)
    s2, g2 = entropy.calculate_entropy(eng, pos, mol["masses"], solvent_id=3, dielectric=47.24, grad_tol=5.0)
    np.testing.assert_array_equal(s, s2)
