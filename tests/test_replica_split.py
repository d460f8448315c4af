"""Small batches: a replica is split into sub-units (disjoint ranges of target atoms, Ctx::split) so that its edge
tiles are spread over several warp groups.  Checked against the CPU oracle and against the unsplit evaluation
(GNNIS_SPLIT=1) of the same inputs; the partial Qbar tables are summed in a fixed order, so results are reproducible
bit for bit."""
import numpy as np
import pytest
import torch

from conftest import rel_rms

pytestmark = pytest.mark.gpu
E_RTOL, F_RRMS = 1e-4, 1e-3


def _run(params, sd, pos, sids, diel, monkeypatch, split):
    from gnnimplicitsolvent_b200.engine import Engine
    if split is None:
        monkeypatch.delenv("GNNIS_SPLIT", raising=False)
    else:
        monkeypatch.setenv("GNNIS_SPLIT", str(split))
    eng = Engine(params, sd)
    eng.set_replicas(pos.shape[0], sids, diel)
    p = torch.from_numpy(pos).cuda()
    e, f = eng.energy_force(p)
    e2, f2 = eng.energy_force(p)
    assert torch.equal(e, e2) and torch.equal(f, f2)          # deterministic
    edges = eng.last_edge_count()
    eng.close()
    return e.cpu().numpy(), f.cpu().numpy(), edges


@pytest.mark.parametrize("N,R", [(50, 1), (37, 3), (100, 2), (200, 1), (24, 5)])
def test_split_vs_oracle_and_unsplit(N, R, monkeypatch):
    from gnnimplicitsolvent_b200 import synth
    from oracle import gnn_oracle as O
    mol = synth.synth_molecule(N, seed=N + 1)
    sd = O.random_state_dict(seed=N)
    pos = synth.conformers(mol["pos"], R, 0.02, seed=R).astype(np.float32)
    sids = [(3 * r + 1) % 39 for r in range(R)]
    diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
    o = O.evaluate(mol["params"], pos, sids, diel, sd)
    e1, f1, n1 = _run(mol["params"], sd, pos, sids, diel, monkeypatch, 1)
    for split in (None, 2, 16):
        e, f, n = _run(mol["params"], sd, pos, sids, diel, monkeypatch, split)
        assert n == n1
        np.testing.assert_allclose(e, o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
        assert rel_rms(f, o["forces"].numpy()) < F_RRMS
        np.testing.assert_allclose(e, e1, rtol=2e-6, atol=1e-4)      # only the summation order differs
        assert rel_rms(f, f1) < 2e-5


def test_split_with_parts_without_edges(monkeypatch):
    """The first half of the atoms forms a normal molecule, the second half is a line of isolated atoms: whole parts
    of the split replica have no GNN edge (empty sub-units next to dense ones)."""
    from gnnimplicitsolvent_b200 import synth
    from oracle import gnn_oracle as O
    N = 48
    mol = synth.synth_molecule(N, seed=9)
    sd = O.random_state_dict(seed=9)
    pos0 = mol["pos"].astype(np.float32).copy()
    pos0[N // 2:] = np.array([[3.0 + 0.9 * i, 0.1 * (i % 2), -2.0] for i in range(N - N // 2)], np.float32)
    pos = np.stack([pos0, pos0 + 0.01]).astype(np.float32)
    o = O.evaluate(mol["params"], pos, [0, 5], [78.5, 24.85], sd)
    for split in (1, None, 16):
        e, f, n = _run(mol["params"], sd, pos, [0, 5], [78.5, 24.85], monkeypatch, split)
        np.testing.assert_allclose(e, o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
        assert rel_rms(f, o["forces"].numpy()) < F_RRMS


@pytest.mark.parametrize("N,R", [(64, 4), (65, 4), (96, 3), (119, 2), (120, 2), (121, 2), (128, 2), (50, 148), (50, 149), (50, 295), (50, 297)])
def test_table_mode_and_split_boundaries(N, R, monkeypatch):
    """Molecule sizes around the shared-memory table modes (two tables up to 64 atoms, one up to ~120, none beyond) and
    batch sizes around the split rule (296 warp groups): the tensor-core path against the independent CUDA-core
    kernels (which neither split nor use those table modes) on the same inputs."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    monkeypatch.delenv("GNNIS_SPLIT", raising=False)
    mol = synth.synth_molecule(N, seed=N + 7)
    sd = O.random_state_dict(seed=N + 3)
    # jitter 0.01 nm: larger displacements of these big synthetic molecules create contacts below 0.05 nm, where the
    # reference's autograd (and therefore the oracle) returns NaN forces
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=R + 1).astype(np.float32)).cuda()
    sids = [(5 * r + 2) % 39 for r in range(R)]
    diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
    eng = Engine(mol["params"], sd)
    eng.set_replicas(R, sids, diel)
    e, f = eng.energy_force(pos)
    e_cc, f_cc = eng.energy_force(pos, cuda_cores=True)
    np.testing.assert_allclose(e.cpu().numpy(), e_cc.cpu().numpy(), rtol=2e-5, atol=1e-3)
    assert rel_rms(f.cpu().numpy(), f_cc.cpu().numpy()) < 5e-4      # two fp32 implementations, both inside the 1e-3 gate
    k = 2
    o = O.evaluate(mol["params"], pos[:k].cpu().numpy(), sids[:k], diel[:k], sd)
    assert bool(torch.isfinite(o["forces"]).all()) and bool(torch.isfinite(f).all())
    np.testing.assert_allclose(e[:k].cpu().numpy(), o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
    assert rel_rms(f[:k].cpu().numpy(), o["forces"].numpy()) < F_RRMS


@pytest.mark.parametrize("N,R", [(50, 297), (50, 300), (50, 433), (50, 444), (50, 1024), (100, 592 + 37), (37, 296 + 150),
                                 (20, 300), (64, 599), (65, 310), (130, 301), (12, 299)])
def test_tail_split_vs_unsplit_and_oracle(N, R, monkeypatch):
    """Batches of at least one full wave (296 warp groups): only the replicas of the last, partly filled wave are split
    (Ctx::split_from).  Same inputs with GNNIS_TAIL_SPLIT=0 (no split at all) -- only the summation order of the split
    replicas may differ --, the first / boundary / last replicas against the CPU oracle, bitwise reproducibility."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    monkeypatch.delenv("GNNIS_SPLIT", raising=False)
    mol = synth.synth_molecule(N, seed=N + 11)
    sd = O.random_state_dict(seed=N + 5)
    pos_h = synth.conformers(mol["pos"], R, 0.01, seed=R + 3).astype(np.float32)
    pos = torch.from_numpy(pos_h).cuda()
    sids = [(7 * r + 1) % 39 for r in range(R)]
    diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
    out = {}
    for tail in ("0", "1"):
        monkeypatch.setenv("GNNIS_TAIL_SPLIT", tail)
        eng = Engine(mol["params"], sd)
        eng.set_replicas(R, sids, diel)
        desc = eng.describe()
        e, f = eng.energy_force(pos)
        e2, f2 = eng.energy_force(pos)
        e3, f3 = eng.energy_force(pos)                      # third call: CUDA-graph replay
        assert torch.equal(e, e2) and torch.equal(f, f2) and torch.equal(e, e3) and torch.equal(f, f3)
        out[tail] = (e.cpu().numpy(), f.cpu().numpy(), desc)
        eng.close()
    assert "split=1" in out["0"][2]
    rem = R % 296
    first = R - rem
    if min(296 // rem, N // 6) >= 2:
        assert f"only)" in out["1"][2] and f"replicas {first}.." in out["1"][2], out["1"][2]
        # whole replicas ahead of the tail are computed exactly as without the split
        assert np.array_equal(out["0"][0][:first], out["1"][0][:first]) and np.array_equal(out["0"][1][:first], out["1"][1][:first])
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=2e-6, atol=1e-4)
    assert rel_rms(out["1"][1], out["0"][1]) < 2e-5
    idx = sorted({0, first - 1, first, min(first + 1, R - 1), R - 1})
    o = O.evaluate(mol["params"], pos_h[idx], [sids[i] for i in idx], [diel[i] for i in idx], sd)
    np.testing.assert_allclose(out["1"][0][idx], o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
    assert rel_rms(out["1"][1][idx], o["forces"].numpy()) < F_RRMS
