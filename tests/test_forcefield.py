"""Vacuum molecular-mechanics terms: CUDA kernel vs the CPU oracle (oracle/mm_oracle.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_rms


def _case(N, R, seed, sigma=0.005):
    from gnnimplicitsolvent_b200 import synth
    mol = synth.synth_molecule(N, seed=seed)
    ff = synth.synth_forcefield(mol)
    pos = synth.conformers(mol["pos"], R, sigma, seed=seed + 3)
    return mol, ff, pos


def test_oracle_forces_match_finite_differences():
    """fp64 oracle: autograd forces == central differences of the energy, term by term present in the total."""
    from oracle import mm_oracle as M
    mol, ff, pos = _case(17, 1, seed=5)
    o = M.evaluate(ff, pos)
    h = 1e-6
    for (a, k) in [(0, 0), (7, 1), (16, 2)]:
        p1, p2 = pos.astype(np.float64).copy(), pos.astype(np.float64).copy()
        p1[0, a, k] += h
        p2[0, a, k] -= h
        fd = -(float(M.evaluate(ff, p1, forces=False)["energy_rep"][0]) - float(M.evaluate(ff, p2, forces=False)["energy_rep"][0])) / (2 * h)
        assert abs(fd - float(o["forces"][0, a, k])) < 1e-4 * max(1.0, abs(fd))


def test_synthetic_forcefield_topology():
    """1-2 / 1-3 pairs are plain exclusions, 1-4 pairs scaled exceptions, equilibrium geometry has zero bonded strain."""
    from oracle import mm_oracle as M
    mol, ff, _ = _case(30, 1, seed=2)
    x = ff["exceptions"]
    zero = (x["chargeprod"] == 0) & (x["epsilon"] == 0)
    assert zero.sum() >= len(ff["bonds"]["idx"]) and (~zero).sum() > 0
    o = M.evaluate(ff, mol["pos"][None], forces=False)
    assert float(o["parts"]["bond"][0]) < 1e-8
    assert float(np.degrees(ff["angles"]["theta0"]).max()) <= 150.0 + 1e-4    # away from the 180 deg singularity


@pytest.mark.gpu
@pytest.mark.parametrize("N,R", [(13, 5), (50, 16), (100, 3), (200, 2)])
def test_vacuum_terms_vs_oracle(N, R):
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import mm_oracle as M
    mol, ff, pos = _case(N, R, seed=N)
    eng = Engine(mol["params"], None)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    eng.set_forcefield(ff)
    e, f = eng.energy_force(torch.from_numpy(pos).cuda(), vacuum_only=True)
    o = M.evaluate(ff, pos)
    scale = float(sum(v.abs().max() for v in o["parts"].values()))     # terms cancel: compare against their size
    assert float((e.cpu().double() - o["energy_rep"]).abs().max()) < 2e-6 * scale
    assert rel_rms(f.cpu().numpy(), o["forces"].numpy()) < 2e-5
    # bit-reproducible
    e2, f2 = eng.energy_force(torch.from_numpy(pos).cuda(), vacuum_only=True)
    assert torch.equal(e, e2) and torch.equal(f, f2)
    # energy only
    e3, _ = eng.energy_force(torch.from_numpy(pos).cuda(), vacuum_only=True, forces=False)
    assert torch.equal(e, e3)


@pytest.mark.gpu
def test_full_hamiltonian_is_gnn_plus_vacuum(weights):
    """GNN + force field = sum of the two separate evaluations; removing the force field restores the GNN alone."""
    from gnnimplicitsolvent_b200.engine import Engine
    mol, ff, pos = _case(50, 8, seed=50)
    eng = Engine(mol["params"], weights)
    eng.set_replicas(8, [0] * 8, [78.5] * 8)
    p = torch.from_numpy(pos).cuda()
    e_gnn, f_gnn = eng.energy_force(p)
    eng.set_forcefield(ff)
    e_mm, f_mm = eng.energy_force(p, vacuum_only=True)
    e_all, f_all = eng.energy_force(p)
    np.testing.assert_allclose(e_all.cpu().numpy(), (e_gnn + e_mm).cpu().numpy(), rtol=1e-5, atol=1e-2)
    assert rel_rms(f_all.cpu().numpy(), (f_gnn + f_mm).cpu().numpy()) < 1e-6
    eng.set_forcefield(None)
    e_back, f_back = eng.energy_force(p)
    assert torch.equal(e_back, e_gnn) and torch.equal(f_back, f_gnn)


@pytest.mark.gpu
def test_minimize_with_forcefield(weights):
    """Batched L-BFGS on the full Hamiltonian (vacuum force field + GNN implicit solvent)."""
    from gnnimplicitsolvent_b200.engine import Engine
    mol, ff, pos = _case(50, 64, seed=50, sigma=0.02)
    eng = Engine(mol["params"], weights)
    eng.set_replicas(64, [0] * 64, [78.5] * 64)
    eng.set_forcefield(ff)
    p0 = torch.from_numpy(pos).cuda()
    e0, f0 = eng.energy_force(p0)
    p1, e1, status, rounds = eng.minimize(p0, grad_tol=10.0, max_iter=500)
    _, f1 = eng.energy_force(p1)
    assert bool((e1 < e0).all()) and int((status == 2).sum()) == 0
    g0 = f0.reshape(64, -1).pow(2).mean(1).sqrt()
    g1 = f1.reshape(64, -1).pow(2).mean(1).sqrt()
    # status 0 = gradient RMS below the threshold; status 3 = the line search can no longer lower the energy: the model's
    # energy surface jumps by up to a few kJ/mol wherever a pair crosses the 0.6 nm graph cutoff (messages do not vanish
    # there, in the reference as here), and a replica whose descent direction crosses such a ledge upwards stops on it
    assert bool((g1[status == 0] < 10.0).all())
    assert int(((status == 0) | (status == 3)).sum()) >= 56
    assert float(g1.median()) < 0.05 * float(g0.median()) and float(g0.min()) > 1000.0
    assert float((e0 - e1).min()) > 100.0
