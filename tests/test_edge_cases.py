"""Edge cases of the hot path on the GPU: empty graphs, minimum sizes, complete graphs, non-finite inputs,
argument errors.  Everything is compared with the CPU oracle on the same inputs."""
import numpy as np
import pytest
import torch

from conftest import rel_rms

pytestmark = pytest.mark.gpu
E_RTOL, F_RRMS = 1e-4, 1e-3


def _setup(N, seed):
    from gnnimplicitsolvent_b200 import synth
    from oracle import gnn_oracle as O
    mol = synth.synth_molecule(N, seed=seed)
    return mol, O.random_state_dict(seed=seed), O


def _compare(params, pos, sd, O, sids=None, diel=None):
    from gnnimplicitsolvent_b200.engine import Engine
    R = pos.shape[0]
    sids = sids or [0] * R
    diel = diel or [78.5] * R
    o = O.evaluate(params, pos, sids, diel, sd)
    eng = Engine(params, sd)
    eng.set_replicas(R, sids, diel)
    e, f = eng.energy_force(torch.from_numpy(pos).cuda())
    np.testing.assert_allclose(e.cpu().numpy(), o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
    assert rel_rms(f.cpu().numpy(), o["forces"].numpy()) < F_RRMS
    return eng, e, f, o


def test_graph_without_edges():
    """All atoms farther apart than the 0.6 nm graph cutoff: E = 0, the message-passing kernels see empty units and
    only the GB-Neck2 terms (all pairs) and the node MLPs contribute."""
    mol, sd, O = _setup(6, 11)
    base = np.array([[0.9 * i, 0.05 * (i % 2), 0.0] for i in range(6)], np.float32)
    pos = np.stack([base, base + 0.01, base * 1.1]).astype(np.float32)
    eng, e, f, o = _compare(mol["params"], pos, sd, O)
    assert eng.last_edge_count() == 0
    row_ptr, col, r = eng.build_neighbors(torch.from_numpy(pos).cuda())
    assert int(row_ptr[-1]) == 0 and col.numel() == 0


def test_mixed_empty_and_dense_replicas():
    """Replica 0 has no edge, replica 1 is the normal molecule, replica 2 has no edge again (empty units between
    non-empty ones, ragged CSR)."""
    mol, sd, O = _setup(12, 5)
    far = np.array([[0.8 * i, 0.0, 0.1 * (i % 3)] for i in range(12)], np.float32)
    pos = np.stack([far, mol["pos"].astype(np.float32), far + 0.02]).astype(np.float32)
    eng, e, f, o = _compare(mol["params"], pos, sd, O, sids=[0, 3, 7], diel=[78.5, 47.24, 32.6])
    rp, cl, rr = O.neighbor_list_numpy(pos)
    row_ptr, col, r = eng.build_neighbors(torch.from_numpy(pos).cuda())
    np.testing.assert_array_equal(row_ptr.cpu().numpy(), rp)
    np.testing.assert_array_equal(col.cpu().numpy(), cl)
    assert rp[12] == 0 and rp[-1] == rp[24]


def test_two_atoms_one_replica():
    mol, sd, O = _setup(2, 3)
    pos = np.array([[[0.0, 0.0, 0.0], [0.15, 0.02, -0.01]]], np.float32)
    eng, e, f, o = _compare(mol["params"], pos, sd, O)
    assert eng.last_edge_count() == 2


def test_complete_graph():
    """Every pair inside the cutoff (degree N - 1 for every atom): the largest rows the CSR builder can produce."""
    N = 27
    mol, sd, O = _setup(N, 9)
    g = np.random.default_rng(1)
    # jittered 3 x 3 x 3 lattice, 0.16 nm spacing: separations between 0.15 and 0.57 nm, all inside the cutoff
    grid = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], np.float32) * 0.16
    pos = np.stack([grid + g.normal(scale=0.003, size=grid.shape) for _ in range(5)]).astype(np.float32)
    assert np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1).max() < 0.59
    eng, e, f, o = _compare(mol["params"], pos, sd, O)
    assert eng.last_edge_count() == 5 * N * (N - 1)


def test_nan_replica_is_isolated():
    """A replica with non-finite coordinates comes back non-finite (the reference propagates NaN the same way,
    `get_energies` turns failures into NaN) and does not disturb the other replicas or hang the kernels."""
    mol, sd, O = _setup(20, 2)
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    pos = synth.conformers(mol["pos"], 4, 0.01, seed=1)
    eng = Engine(mol["params"], sd)
    eng.set_replicas(4, [0] * 4, [78.5] * 4)
    e0, f0 = eng.energy_force(torch.from_numpy(pos).cuda())
    bad = pos.copy()
    bad[2, 5] = np.nan
    e1, f1 = eng.energy_force(torch.from_numpy(bad).cuda())
    torch.cuda.synchronize()
    keep = [0, 1, 3]
    np.testing.assert_array_equal(e1.cpu().numpy()[keep], e0.cpu().numpy()[keep])
    np.testing.assert_array_equal(f1.cpu().numpy()[keep], f0.cpu().numpy()[keep])
    assert not np.isfinite(e1.cpu().numpy()[2])


def test_argument_errors_do_not_crash():
    from gnnimplicitsolvent_b200.engine import Engine, GnnisError
    mol, sd, O = _setup(8, 4)
    eng = Engine(mol["params"], sd)
    with pytest.raises((GnnisError, AssertionError, ValueError)):
        eng.set_replicas(3, [0, 0], [78.5, 78.5, 78.5])          # lengths disagree
    with pytest.raises((GnnisError, AssertionError, ValueError)):
        eng.set_replicas(2, [0, 99], [78.5, 78.5])               # solvent id outside the embedding table
    eng.set_replicas(2, [0, 1], [78.5, 47.24])
    with pytest.raises((GnnisError, AssertionError, ValueError, RuntimeError)):
        eng.energy_force(torch.zeros(3, 8, 3, device="cuda"))    # wrong replica count


@pytest.mark.gpu
def test_size_limit_and_path_description(weights):
    """No silent degradation: more than 1024 atoms per replica is refused with a message that names the limit, and
    gnnis_describe tells which path a set-up runs on (stash / recomputing adjoint, tcgen05 / CUDA cores)."""
    from gnnimplicitsolvent_b200.engine import Engine, GnnisError
    from gnnimplicitsolvent_b200 import synth
    big = np.zeros((1025, 7), dtype=np.float32)
    big[:, 1] = 0.15
    with pytest.raises(GnnisError, match="at most 1024"):
        Engine(big, weights)
    mol = synth.synth_molecule(50, seed=50)
    eng = Engine(mol["params"], weights)
    eng.set_replicas(64, [0] * 64, [78.5] * 64)
    d = eng.describe()
    assert "N=50" in d and "edge_kernels=tcgen05" in d and "adjoint=stash" in d
    mol = synth.synth_molecule(100, seed=100)
    eng = Engine(mol["params"], weights)
    eng.set_replicas(8, [0] * 8, [78.5] * 8)
    assert "edge_kernels=tcgen05" in eng.describe()
