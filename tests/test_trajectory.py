"""Trajectory I/O and frame-energy read-back: file layout on CPU, the simulation driver on GPU."""
import os

import numpy as np
import pytest
import torch

from gnnimplicitsolvent_b200 import trajectory


def test_writer_layout_replica_as_chain(tmp_path):
    R, N, F = 3, 5, 4
    rng = np.random.default_rng(0)
    frames = rng.normal(size=(F, R, N, 3)).astype(np.float32)
    prefix = str(tmp_path / "mol_GNN_0_42")
    with trajectory.TrajectoryWriter(prefix, F, R, N, elements=["C", "H", "H", "O", "N"]) as w:
        for f in range(F):
            w.append(frames[f], time_ps=0.2 * (f + 1), step=100 * (f + 1), potential=np.arange(R) + f,
                     temperature=np.full(R, 300.0 + f))
    xyz, meta = trajectory.load_trajectory(prefix)
    assert xyz.shape == (F, R * N, 3) and xyz.dtype == np.float32
    np.testing.assert_array_equal(np.asarray(xyz).reshape(F, R, N, 3), frames)
    np.testing.assert_array_equal(meta["chainid"], np.repeat(np.arange(R), N))
    np.testing.assert_allclose(meta["time"], [0.2, 0.4, 0.6, 0.8])
    assert meta["potentialEnergy"].shape == (F, R) and list(meta["element"]) == ["C", "H", "H", "O", "N"]
    # the reference's reader: one trajectory per chain
    per_rep = trajectory.load_parallel_traj(prefix, R)
    assert len(per_rep) == R
    for r in range(R):
        np.testing.assert_array_equal(per_rep[r], frames[:, r])
    with pytest.raises(ValueError):
        trajectory.load_parallel_traj(prefix, R + 1)
    log = open(prefix + "_log.txt").read().strip().split("\n")
    assert log[0].startswith('#"Step"') and len(log) == F + 1 and log[1].split(",")[0] == "100"


def test_writer_refuses_overflow_and_partial_close(tmp_path):
    prefix = str(tmp_path / "t")
    w = trajectory.TrajectoryWriter(prefix, 3, 1, 2)
    w.append(np.zeros((1, 2, 3), np.float32), 0.1, 1)
    w.close()
    xyz, meta = trajectory.load_trajectory(prefix)
    assert xyz.shape[0] == 1 and int(meta["n_frames"]) == 1      # only the frames that were written are exposed
    w2 = trajectory.TrajectoryWriter(prefix, 1, 1, 2)
    w2.append(np.zeros((1, 2, 3), np.float32), 0.1, 1)
    with pytest.raises(IndexError):
        w2.append(np.zeros((1, 2, 3), np.float32), 0.2, 2)
    w2.close()


def test_kinetic_temperature_equipartition():
    g = torch.Generator().manual_seed(1)
    R, N, T = 64, 200, 300.0
    m = np.linspace(1.0, 16.0, N).astype(np.float32)
    sigma = torch.sqrt(torch.tensor(trajectory.KB_KJ_MOL_K * T) / torch.from_numpy(m)).reshape(1, N, 1)
    v = torch.randn(R, N, 3, generator=g) * sigma
    t = trajectory.kinetic_temperature(v, m)
    assert t.shape == (R,)
    assert abs(float(t.mean()) - T) < 3.0


@pytest.mark.gpu
def test_run_simulation_and_get_energies(tmp_path):
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "gnn_weights.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    N, R = 24, 16
    mol = synth.synth_molecule(N, seed=5)
    ff = synth.synth_forcefield(mol)
    eng = Engine(mol["params"], sd)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    eng.set_forcefield(ff)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.005, seed=2)).cuda()
    vel = torch.zeros_like(pos)
    m = np.asarray(mol["masses"], np.float32)
    prefix = str(tmp_path / "run")
    trajectory.run_simulation(eng, pos, vel, m, n_steps=60, n_interval=20, minimize=True, prefix=prefix, dt=0.0005,
                              friction=5.0, seed=3, elements=mol["elements"])
    xyz, meta = trajectory.load_trajectory(prefix)
    assert xyz.shape == (3, R * N, 3) and np.isfinite(np.asarray(xyz)).all()
    np.testing.assert_allclose(meta["time"], [0.01, 0.02, 0.03], rtol=1e-6)
    # the last stored frame is the final state, and the stored energies are those of the stored frames
    np.testing.assert_array_equal(np.asarray(xyz[-1]).reshape(R, N, 3), pos.cpu().numpy())
    e_last, _ = eng.energy_force(pos, forces=False)
    np.testing.assert_allclose(meta["potentialEnergy"][-1], e_last.double().cpu().numpy(), rtol=1e-6)
    assert np.isfinite(meta["temperature"]).all() and (meta["temperature"] > 0).all()
    # frame energies of replica 0, batched (the reference re-evaluates frame by frame on a one-replica system)
    rep0 = trajectory.load_parallel_traj(prefix, 1)[0]
    e_frames = trajectory.get_energies(eng, rep0)
    np.testing.assert_allclose(e_frames, meta["potentialEnergy"][:, 0], rtol=2e-5, atol=1e-3)
