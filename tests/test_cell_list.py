"""Cell-list back-end of the neighbour builder (csrc/cell_list.cu; `north_star`'s "cell-list radius-graph builder emitting
sorted CSR edges", replacing torch_cluster.radius_graph, GNN_Models.py:345-357 / 887-905): bit-identical to the all-pairs
back-end, to the reference's golden edge lists and to the numpy oracle, for both predicates, compact and extended
geometries (one cell ... 8 x 8 x 8 cells and beyond), and as the graph under a full evaluation."""
import numpy as np
import pytest
import torch

from conftest import CASES, load_case

pytestmark = pytest.mark.gpu


def _lists(eng, pos, predicate="run"):
    p = torch.from_numpy(pos).cuda()
    a = eng.build_neighbors(p, predicate=predicate, backend="allpairs")
    b = eng.build_neighbors(p, predicate=predicate, backend="cell")
    return [x.cpu().numpy() for x in a], [x.cpu().numpy() for x in b]


def _assert_same(a, b):
    for x, y, name in zip(a, b, ("row_ptr", "col", "r")):
        assert x.dtype == y.dtype and x.shape == y.shape, name
        np.testing.assert_array_equal(x.view(np.int32), y.view(np.int32), err_msg=name)   # bytes, not values


@pytest.mark.parametrize("name", CASES)
def test_cell_list_vs_reference_golden(name, weights):
    """The edge lists and fp32 distances the unmodified reference produced (tests/golden), through the cell list."""
    from gnnimplicitsolvent_b200.engine import Engine
    c = load_case(name)
    eng = Engine(c["params"], weights)
    R, N, _ = c["pos"].shape
    eng.set_replicas(R, c["solvent_ids"], c["dielectrics"])
    row_ptr, col, r = eng.build_neighbors(torch.from_numpy(c["pos"]).cuda(), backend="cell")
    tgt = np.repeat(np.arange(R * N), np.diff(row_ptr.cpu().numpy()))
    np.testing.assert_array_equal(col.cpu().numpy(), c["edge_src"])
    np.testing.assert_array_equal(tgt, c["edge_tgt"])
    np.testing.assert_array_equal(r.cpu().numpy(), c["edge_r"])


def _params(N):
    """[N,7] table for neighbour-list tests: the rows of a 50-atom synthetic molecule, repeated (building a self-avoiding
    1024-atom molecule takes minutes on the host and the graph does not depend on the parameters)."""
    from gnnimplicitsolvent_b200 import synth
    return np.resize(synth.synth_molecule(50, seed=50)["params"], (N, 7))


def _chain(N, step, seed, wiggle=0.08):
    """Self-avoiding-ish random walk: extended geometries spanning many cells."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(N, 3))
    d[:, 0] += 1.5                       # drift along x
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x = np.cumsum(d * step, axis=0) + rng.normal(scale=wiggle, size=(N, 3))
    return x.astype(np.float32)


@pytest.mark.parametrize("N,R,step", [(2, 5, 0.3), (13, 9, 0.15), (50, 33, 0.12), (200, 6, 0.14), (333, 3, 0.25),
                                      (1024, 2, 0.05), (1024, 2, 0.3), (700, 2, 2.0)])
@pytest.mark.parametrize("predicate", ["run", "data"])
def test_cell_list_vs_all_pairs_and_oracle(N, R, step, predicate):
    """Compact blobs (one cell), chains spanning more than 8 cells per dimension (cells wider than the cutoff), sparse chains
    without any edge; R not a multiple of anything; both predicates.  The two back-ends must agree to the byte, and with
    the numpy oracle where that finishes quickly."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    pos = np.stack([_chain(N, step, 1000 * N + r) for r in range(R)])
    eng = Engine(_params(N), None)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    a, b = _lists(eng, pos, predicate)
    _assert_same(a, b)
    if N <= 333:
        rp, cl, rr = O.neighbor_list_numpy(pos, predicate=predicate)
        np.testing.assert_array_equal(b[0], rp)
        np.testing.assert_array_equal(b[1], cl)


def test_cell_list_degenerate_boxes():
    """Planar, linear and single-point extents (a zero-width bounding box in one, two or three dimensions), atoms exactly
    on cell faces, and a pair inside the 1e-6 eps band of the cutoff in different cells."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    N = 24
    k = np.arange(N, dtype=np.float32)
    line = np.stack([0.3003 * k, np.zeros(N), np.zeros(N)], 1)                       # multiples of half a cell
    plane = np.stack([0.6006 * (k % 6), 0.6006 * (k // 6), np.zeros(N)], 1)          # atoms on cell corners
    point = np.zeros((N, 3), np.float32) + 0.25
    band = line.copy()
    band[1] = band[0] + np.array([0.6 - 5e-7, 0.0, 0.0], np.float32)                 # straddles the eps band
    band[2] = band[0] + np.array([0.6 + 5e-7, 0.0, 0.0], np.float32)
    pos = np.stack([line, plane, point, band]).astype(np.float32)
    eng = Engine(_params(N), None)
    eng.set_replicas(4, [0] * 4, [78.5] * 4)
    for predicate in ("run", "data"):
        a, b = _lists(eng, pos, predicate)
        _assert_same(a, b)
    assert np.diff(b[0])[2 * N:3 * N].tolist() == [N - 1] * N                        # coincident atoms: complete graph


def test_evaluation_on_cell_list_graph_is_bit_identical(weights):
    """A whole evaluation (energies, forces) on the graph of either back-end: same bits (the CSR arrays are the same, the
    rest of the pipeline is unchanged), including the data-path predicate and a second call (CUDA-graph replay)."""
    from gnnimplicitsolvent_b200.engine import Engine
    c = load_case("n100_mixed")
    eng = Engine(c["params"], weights)
    R = c["pos"].shape[0]
    eng.set_replicas(R, c["solvent_ids"], c["dielectrics"])
    pos = torch.from_numpy(c["pos"]).cuda()
    for kw in ({}, {"data_predicate": True}):
        from gnnimplicitsolvent_b200 import _lib
        e0, f0 = eng.energy_force(pos, extra_flags=_lib.FLAG_ALL_PAIRS, **kw)
        for _ in range(3):
            e1, f1 = eng.energy_force(pos, cell_list=True, **kw)
            assert torch.equal(e0, e1) and torch.equal(f0, f1)
    assert "neighbors=" in eng.describe()


def test_cell_list_non_finite_positions():
    """A NaN coordinate and an infinite coordinate: the atom has no edges with either back-end (every predicate involving
    it is false), the bounding box / cell assignment of the others is not disturbed."""
    from gnnimplicitsolvent_b200.engine import Engine
    N, R = 40, 3
    pos = np.stack([_chain(N, 0.12, 77 + r) for r in range(R)])
    pos[0, 7, 1] = np.nan
    pos[1, 0, 2] = np.inf
    pos[2, 39, 0] = -np.inf
    eng = Engine(_params(N), None)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    for predicate in ("run", "data"):
        a, b = _lists(eng, pos, predicate)
        _assert_same(a, b)
        deg = np.diff(b[0]).reshape(R, N)
        assert deg[0, 7] == 0 and deg[1, 0] == 0 and deg[2, 39] == 0 and deg.sum() > 0
        assert not np.isin(b[1], [7, N + 0, 2 * N + 39]).any()
