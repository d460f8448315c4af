"""Batched L-BFGS minimiser and LangevinMiddle integrator.  OpenMM's own
LocalEnergyMinimizer / LangevinMiddleIntegrator cannot run here (parity unpinned), so
correctness is defined physically: monotone decrease to the gradient threshold, agreement of the minima with
scipy's BFGS on the CPU oracle's potential, equipartition / harmonic variance, bit-reproducibility per seed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _system(N, R, seed, sigma=0.02):
    from gnnimplicitsolvent_b200 import synth
    mol = synth.synth_molecule(N, seed=seed)
    pos = synth.conformers(mol["pos"], R, sigma, seed=seed + 1)
    # Stand-in for the vacuum force field (SURVEY.md N1 is not built yet): an elastic network over ALL atom pairs
    # at the base geometry.  It keeps the implicit-solvent potential bounded (GB alone has no repulsion) and gives
    # every replica one well-defined basin, so minima found by different optimisers can be compared.
    iu = np.triu_indices(N, 1)
    bonds = np.stack(iu, axis=1).astype(np.int32)
    r0 = np.linalg.norm(mol["pos"][bonds[:, 0]] - mol["pos"][bonds[:, 1]], axis=1).astype(np.float32)
    k = np.full(len(bonds), 2.0e4, dtype=np.float32)      # kJ/mol/nm^2
    return mol, pos, bonds, r0, k


def test_minimizer_reaches_gradient_threshold(weights):
    from gnnimplicitsolvent_b200.engine import Engine
    mol, pos, bonds, r0, k = _system(21, 16, seed=4)
    eng = Engine(mol["params"], weights)
    eng.set_replicas(16, [0] * 16, [78.5] * 16)
    eng.set_bonds(bonds, r0, k)
    p0 = torch.from_numpy(pos).cuda()
    e0, f0 = eng.energy_force(p0)
    p1, e1, status, rounds = eng.minimize(p0, grad_tol=5.0, max_iter=400)
    e_chk, f_chk = eng.energy_force(p1)
    assert rounds > 0 and int((status == 2).sum()) == 0
    assert bool((e1 <= e0 + 1e-3).all())                                   # never worse than the start
    np.testing.assert_allclose(e1.cpu().numpy(), e_chk.cpu().numpy(), rtol=1e-5, atol=1e-2)
    grms = f_chk.reshape(16, -1).pow(2).mean(1).sqrt()
    conv = status == 0
    assert bool((grms[conv] < 5.0).all())                                  # status 0 means below the stated threshold
    assert int(((status == 0) | (status == 3)).sum()) >= 14                # the rest ended on a line-search stall ...
    assert float(grms[status == 3].max() if bool((status == 3).any()) else 0.0) < 100.0   # ... not far from a minimum
    assert float(f0.reshape(16, -1).pow(2).mean(1).sqrt().min()) > 50.0    # the start was far from a minimum


def test_minimizer_agrees_with_scipy_on_oracle_potential(weights):
    """GB-only potential + harmonic bonds, one replica: the batched GPU L-BFGS and scipy's BFGS on the CPU
    oracle find the same minimum energy."""
    import scipy.optimize
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    mol, pos, bonds, r0, k = _system(13, 1, seed=9, sigma=0.01)
    N = 13
    x = torch.tensor(mol["params"], dtype=torch.float64)
    ei = O.all_pairs_index(N, 1)
    d0, m0 = O.neck_tables()
    bi = torch.from_numpy(bonds.astype(np.int64))

    def fun(v):
        p = torch.tensor(v.reshape(N, 3), dtype=torch.float64, requires_grad=True)
        r = O.eps_distance(p, ei[0], ei[1])
        B, _ = O.born_radii(x, ei[0], ei[1], r, d0.double(), m0.double())
        e = O.gb_energy(torch.stack((B, x[:, 0]), 1), ei[0], ei[1], r, torch.full((N, 1), 78.5, dtype=torch.float64)).sum()
        bl = (p[bi[:, 0]] - p[bi[:, 1]]).norm(dim=1)
        e = e + (0.5 * torch.from_numpy(k).double() * (bl - torch.from_numpy(r0).double()) ** 2).sum()
        (g,) = torch.autograd.grad(e, p)
        return float(e.detach()), g.numpy().ravel().copy()

    res = scipy.optimize.minimize(fun, pos[0].astype(np.float64).ravel(), jac=True, method="BFGS",
                                  options={"maxiter": 2000, "gtol": 1e-3})
    assert res.success
    eng = Engine(mol["params"], None)
    eng.set_replicas(1, [0], [78.5])
    eng.set_bonds(bonds, r0, k)
    p1, e1, status, _ = eng.minimize(torch.from_numpy(pos).cuda(), grad_tol=0.5, max_iter=1000, gb_only=True)
    assert abs(float(e1[0]) - res.fun) < 5e-2 + 1e-4 * abs(res.fun), (float(e1[0]), res.fun)


def test_minimizer_32_replicas_full_model_vs_scipy(weights):
    """The full Hamiltonian (GB-Neck2 + GNN, shipped weights) plus the elastic network, N = 50, 32 replicas.  The model's
    energy surface has ledges at cutoff crossings and several nearby minima, so two optimisers started at the same
    point need not end in the same one (measured: median difference 1.9 kJ/mol, either sign).  What is tested is that
    the batched GPU minimiser ends in genuine minima of the ORACLE's potential: every replica reports status 0 with a
    gradient RMS below the threshold, and scipy's L-BFGS-B on the CPU oracle (all replicas as one separable problem),
    started from the GPU's final positions, cannot lower any replica's energy by more than a fraction of a kJ/mol."""
    import scipy.optimize
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    N, R = 50, 32
    mol, pos, bonds, r0, k = _system(N, R, seed=50, sigma=0.01)
    sids, diel = [0] * R, [78.5] * R
    bi = bonds.astype(np.int64)
    kk, rr0 = k.astype(np.float64), r0.astype(np.float64)
    torch.set_num_threads(8)

    def fun(v):
        p = v.reshape(R, N, 3)
        o = O.evaluate(mol["params"], p.astype(np.float32), sids, diel, weights)
        d = p[:, bi[:, 0]] - p[:, bi[:, 1]]
        bl = np.linalg.norm(d, axis=2)
        e_b = (0.5 * kk * (bl - rr0) ** 2).sum(axis=1)
        gb = (kk * (bl - rr0) / bl)[..., None] * d
        g = -o["forces"].double().numpy()
        np.add.at(g, (slice(None), bi[:, 0]), gb)
        np.add.at(g, (slice(None), bi[:, 1]), -gb)
        e_rep = o["energy_rep"].double().numpy() + e_b
        fun.last = e_rep
        return float(e_rep.sum()), g.ravel()

    eng = Engine(mol["params"], weights)
    eng.set_replicas(R, sids, diel)
    eng.set_bonds(bonds, r0, k)
    e_start, _ = eng.energy_force(torch.from_numpy(pos).cuda())
    p1, e1, status, rounds = eng.minimize(torch.from_numpy(pos).cuda(), grad_tol=2.0, max_iter=800)
    e_gpu = e1.double().cpu().numpy()
    n_evals, grms = eng.minimize_stats()
    st = status.cpu().numpy()
    assert not np.any(st == 2) and np.all(st == 0)
    assert float(grms.max()) < 2.0 and int(n_evals.max()) <= rounds          # status 0 = below the stated threshold
    x0 = p1.double().cpu().numpy().ravel()
    fun(x0)
    e_oracle_at_gpu = fun.last.copy()
    np.testing.assert_allclose(e_oracle_at_gpu, e_gpu, rtol=1e-5, atol=2e-2)    # same potential on both sides
    res = scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B",
                                  options={"maxiter": 60, "maxfun": 80, "gtol": 0.05, "ftol": 1e-14, "maxcor": 10})
    fun(res.x)
    gain = e_oracle_at_gpu - fun.last                                         # what scipy still finds, per replica
    print(f"minimiser: rounds {rounds}, evaluations per replica median {int(n_evals.median())}, start -> end "
          f"{float((e_start.double().cpu().numpy() - e_gpu).mean()):.1f} kJ/mol; scipy from the GPU minima gains "
          f"median {np.median(gain):.4f}, max {gain.max():.4f} kJ/mol")
    assert np.median(gain) < 0.05 and gain.max() < 0.5


def test_langevin_reproducible_and_thermalises():
    """Harmonic bonds + GB-only forces: same seed -> bit-identical trajectory; kinetic temperature ~ 300 K."""
    from gnnimplicitsolvent_b200.engine import Engine
    mol, pos, bonds, r0, k = _system(21, 64, seed=2, sigma=0.002)
    eng = Engine(mol["params"], None)
    eng.set_replicas(64, [0] * 64, [78.5] * 64)
    eng.set_bonds(bonds, r0, k)
    m = np.asarray(mol["masses"], dtype=np.float32)

    def run(seed, steps):
        p = torch.from_numpy(pos).cuda().contiguous()
        v = torch.zeros_like(p)
        eng.langevin(p, v, m, steps, temperature=300.0, friction=5.0, dt=0.0005, seed=seed, gb_only=True)
        torch.cuda.synchronize()
        return p, v

    p1, v1 = run(7, 400)
    p2, v2 = run(7, 400)
    assert torch.equal(p1, p2) and torch.equal(v1, v2)
    p3, v3 = run(8, 400)
    assert not torch.equal(v1, v3)
    # equipartition: <m v^2> = kT per degree of freedom (no constraints in this integrator)
    p, v = run(11, 3000)
    mt = torch.from_numpy(m).cuda().reshape(1, -1, 1)
    kT = 0.0083144626 * 300.0
    t_kin = float((mt * v * v).mean() / kT)
    assert 0.85 < t_kin < 1.15, t_kin
    assert bool(torch.isfinite(p).all())



def test_langevin_with_hbond_constraints(weights):
    """constraints=HBonds at dt = 2 fs on the full Hamiltonian (vacuum force field + GNN): X-H distances stay at their
    force-field lengths to the SHAKE tolerance, the kinetic temperature of the 3N - n_c degrees of freedom is ~300 K,
    and the trajectory is bit-reproducible."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    mol = synth.synth_molecule(30, seed=8)
    ff = synth.synth_forcefield(mol)
    R = 64
    pos = synth.conformers(mol["pos"], R, 0.0, seed=1)
    eng = Engine(mol["params"], weights)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    eng.set_forcefield(ff)
    hb = [k for k, (i, j) in enumerate(ff["bonds"]["idx"]) if mol["elements"][i].startswith("H") or mol["elements"][j].startswith("H")]
    cidx, cd = ff["bonds"]["idx"][hb], ff["bonds"]["r0"][hb]
    eng.set_constraints(cidx, cd)
    m = np.asarray(mol["masses"], dtype=np.float32)

    def run(seed, steps):
        p = torch.from_numpy(pos).cuda().contiguous()
        v = torch.zeros_like(p)
        eng.langevin(p, v, m, steps, temperature=300.0, friction=10.0, dt=0.002, seed=seed)
        torch.cuda.synchronize()
        return p, v

    p, v = run(5, 1500)
    assert bool(torch.isfinite(p).all())
    d = (p[:, cidx[:, 0]] - p[:, cidx[:, 1]]).norm(dim=-1).cpu().numpy()
    assert float(np.abs(d / cd[None] - 1.0).max()) < 5e-5
    mt = torch.from_numpy(m).cuda().reshape(1, -1, 1)
    dof = 3 * 30 - len(hb)
    t_kin = float((mt * v * v).sum() / R / (dof * 0.0083144626 * 300.0))
    assert 0.8 < t_kin < 1.2, t_kin
    p2, v2 = run(5, 1500)
    assert torch.equal(p, p2) and torch.equal(v, v2)


def test_steepest_descent_mode_reaches_the_same_minima():
    """history = -1 selects steepest descent (`north_star`: "a batched L-BFGS / steepest-descent minimiser"): same line search
    and stopping rules, no curvature pairs.  On a smooth single-basin potential (GB-Neck2 + elastic network; the GNN term is
    left out because its energy surface has ledges at the graph cutoff, DESIGN.md 3, on which two optimisers need not stop
    at the same one) it must arrive where L-BFGS arrives, just with more evaluations."""
    from gnnimplicitsolvent_b200.engine import Engine
    mol, pos, bonds, r0, k = _system(21, 24, seed=6)
    eng = Engine(mol["params"], None)
    eng.set_replicas(24, [0] * 24, [78.5] * 24)
    eng.set_bonds(bonds, r0, k)
    p0 = torch.from_numpy(pos).cuda()
    e0, _ = eng.energy_force(p0, gb_only=True)
    p_l, e_l, st_l, rounds_l = eng.minimize(p0, grad_tol=2.0, max_iter=600, gb_only=True)
    n_l, _ = eng.minimize_stats()
    p_s, e_s, st_s, rounds_s = eng.minimize(p0, grad_tol=2.0, max_iter=6000, history=-1, gb_only=True)
    n_s, g_s = eng.minimize_stats()
    assert int((st_s == 2).sum()) == 0 and bool((e_s <= e0 + 1e-3).all())
    ok = (st_l == 0) & (st_s == 0)
    assert int(ok.sum()) >= 20, (st_l.tolist(), st_s.tolist())               # both below the threshold
    assert bool((g_s[st_s == 0] < 2.0).all())
    # same minimum: the energies agree to the flatness a 2 kJ/mol/nm gradient threshold leaves
    assert float((e_l[ok] - e_s[ok]).abs().max()) < 0.05
    assert float(n_s[ok].float().median()) > float(n_l[ok].float().median())  # steepest descent needs more evaluations
    e_chk, _ = eng.energy_force(p_s, gb_only=True)
    np.testing.assert_allclose(e_s.cpu().numpy(), e_chk.cpu().numpy(), rtol=1e-5, atol=1e-2)
