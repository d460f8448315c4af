"""The reference's named entry points (Simulation/helper_functions.py, Simulation/Simulator.py, __init__.py:3-8) on
the batched engine: package surface and argument lists (CPU), behaviour against direct Engine calls (GPU)."""
import inspect
import os

import numpy as np
import pytest
import torch

import gnnimplicitsolvent_b200 as G
from gnnimplicitsolvent_b200 import helper_functions as H
from gnnimplicitsolvent_b200 import synth
from gnnimplicitsolvent_b200.Simulator import Molecule, Multi_simulator, Simulator

HERE = os.path.dirname(os.path.abspath(__file__))


def _sd():
    z = np.load(os.path.join(HERE, "golden", "gnn_weights.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_package_exports_the_reference_names():
    # /root/reference/__init__.py:3-8
    for name in ("minimize_mol", "create_gnn_model", "calculate_entropy", "GNN3_Multisolvent_embedding_run_multiple_Delta"):
        assert hasattr(G, name), name
    assert len(G.SOLVENT_DICT) == 39 and G.SOLVENT_DICT["DMSO"] == {"solvent_id": 3, "dielectric": 47.24}
    assert G.SOLVENT_DICT["tip3p"]["solvent_id"] == 0 and G.SOLVENT_DICT["oxylol"]["solvent_id"] == 38


def test_argument_lists_follow_the_reference():
    # helper_functions.py:931-947 (minimize_mol), :1688-1703 (calculate_entropy), :826-838 (create_gnn_model)
    ref_minimize = ["mol", "solvent", "model_path", "solvent_dict", "return_traj", "tolerance", "max_iterations", "gnn_sim",
                    "return_gnn_sim", "strides", "cache", "save_name", "partial_charges", "forcefield", "constraints"]
    assert list(inspect.signature(H.minimize_mol).parameters)[:len(ref_minimize)] == ref_minimize
    d = {k: v.default for k, v in inspect.signature(H.minimize_mol).parameters.items()}
    assert d["tolerance"] == 1e-4 and d["max_iterations"] == 0 and d["strides"] == 1 and d["return_traj"] is False
    ref_entropy = ["mol", "solvent", "model_path", "solvent_dict", "tolerance", "max_iterations", "gnn_sim", "strides",
                   "cache", "save_name", "partial_charges", "forcefield", "return_free_energy", "return_traj"]
    assert list(inspect.signature(H.calculate_entropy).parameters)[:len(ref_entropy)] == ref_entropy
    ref_model = ["mol", "model_class", "cache", "solvent_dict", "solvent", "num_solvents", "model_dict", "partial_charges",
                 "forcefield", "device", "jit"]
    assert list(inspect.signature(H.create_gnn_model).parameters) == ref_model
    # Simulator.py:341-347, :768, :800, :828
    assert list(inspect.signature(Simulator.run_simulation).parameters)[1:] == ["n_steps", "n_interval", "minimize", "workfolder"]
    for m in ("set_positions", "calculate_energy", "calculate_forces", "run_simulation", "hdf5_loc", "log_loc", "pos"):
        assert hasattr(Simulator, m)
    for m in ("setup_replicates", "create_ref_system", "set_random_positions_for_each_replicate", "run_ref_simulation"):
        assert hasattr(Multi_simulator, m)


def test_molecule_record_behaves_like_an_rdkit_mol():
    mol = synth.synth_mol_object(12, 5, seed=2)
    assert mol.GetNumConformers() == 5 and mol.GetNumAtoms() == 12
    p = mol.GetConformer(3).GetPositions()          # Angstrom
    assert p.shape == (12, 3) and np.allclose(p, mol.conformers_nm[3] * 10)
    mol.GetConformer(3).SetAtomPosition(0, [1.0, 2.0, 3.0])
    assert np.allclose(mol.conformers_nm[3][0], [0.1, 0.2, 0.3])
    idx, d = mol.hbond_constraints()
    assert len(idx) == sum(e.startswith("H") for e in mol.elements) and np.all(d > 0.05)
    sim_positions = np.stack([mol.GetConformer(i).GetPositions() / 10 for i in range(2, 4)]).reshape(-1, 3)
    assert sim_positions.shape == (24, 3)


def test_simulator_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gnnimplicitsolvent_b200.engine import GnnisError
    with pytest.raises((GnnisError, OSError, RuntimeError)):
        Simulator(mol=synth.synth_mol_object(10, 1), state_dict=_sd())


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_simulator_energy_and_forces_match_the_engine():
    from gnnimplicitsolvent_b200.engine import Engine
    mol = synth.synth_mol_object(21, 4, seed=5)
    sd = _sd()
    sim = Multi_simulator(mol=mol, state_dict=sd, solvent_model=[3] * 4, solvent_dielectric=[47.24] * 4, num_rep=4)
    sim.setup_replicates()                                   # force self-check of the reference
    pos = mol.conformers_nm.reshape(-1, 3)
    sim.set_positions(pos)
    eng = Engine(mol.parameters, sd)
    eng.set_forcefield(mol.forcefield)
    eng.set_replicas(4, [3] * 4, [47.24] * 4)
    e, f = eng.energy_force(torch.from_numpy(mol.conformers_nm.astype(np.float32)).cuda())
    assert abs(sim.calculate_energy() - float(e.double().sum())) < 1e-3 * abs(float(e.double().sum())) * 1e-3 + 1e-2
    fs = sim.calculate_forces()
    assert fs.shape == (3, 4 * 21)
    assert np.allclose(fs, f.reshape(-1, 3).t().double().cpu().numpy(), rtol=1e-5, atol=1e-3)
    # single-copy reference system: energy of conformer 2 on its own
    sim._ref_system.set_positions(mol.conformers_nm[2])
    assert abs(sim._ref_system.calculate_energy() - float(e[2])) < 2e-3 + 1e-5 * abs(float(e[2]))


@pytest.mark.gpu
def test_minimize_mol_strides_and_energies():
    mol = synth.synth_mol_object(13, 11, seed=7, sigma=0.02)
    start = mol.conformers_nm.copy()
    sd = _sd()
    out = H.minimize_mol(mol, "DMSO", model_path={"model": sd}, return_traj=True, return_gnn_sim=True, strides=3)
    mol2, traj, energies, sim = out
    assert traj.n_frames == 11 and energies.shape == (11,) and np.all(np.isfinite(energies))
    assert sim.minimize_status.shape == (11,)                # 3 blocks of 3 + the 2 left-over conformers
    assert not np.allclose(mol2.conformers_nm, start)        # the molecule carries the minimised coordinates
    assert np.allclose(mol2.conformers_nm, traj.xyz, atol=1e-6)
    # energies are those of the single-copy system at the minimised geometries, and lower than at the start
    from gnnimplicitsolvent_b200.engine import Engine
    eng = Engine(mol.parameters, sd)
    eng.set_forcefield(mol.forcefield)
    eng.set_replicas(11, [3] * 11, [47.24] * 11)
    e_min, f_min = eng.energy_force(torch.from_numpy(traj.xyz).cuda())
    e_start, _ = eng.energy_force(torch.from_numpy(start.astype(np.float32)).cuda())
    assert np.allclose(energies, e_min.double().cpu().numpy(), rtol=1e-5, atol=2e-3)
    assert np.all(e_min.cpu().numpy() < e_start.cpu().numpy())
    grms = f_min.reshape(11, -1).pow(2).mean(dim=1).sqrt().cpu().numpy()
    grms0 = (eng.energy_force(torch.from_numpy(start.astype(np.float32)).cuda())[1]).reshape(11, -1).pow(2).mean(dim=1).sqrt().cpu().numpy()
    assert not np.any(sim.minimize_status == 2)              # nothing went non-finite
    assert np.all(grms[sim.minimize_status == 0] < 1.0)      # status 0 = below the threshold the entry point uses
    assert np.median(grms) < 0.05 * np.median(grms0)         # every conformer came down by orders of magnitude
    # plain call returns (mol, energies)
    mol3, e3 = H.minimize_mol(synth.synth_mol_object(13, 4, seed=7, sigma=0.02), "tip3p", model_path={"model": sd})
    assert e3.shape == (4,)


@pytest.mark.gpu
def test_minimize_mol_vacuum_and_gbneck_branches():
    sd = _sd()
    mol = synth.synth_mol_object(13, 4, seed=9, sigma=0.02)
    _, e_vac = H.minimize_mol(mol, "vac", model_path={"model": sd})
    mol = synth.synth_mol_object(13, 4, seed=9, sigma=0.02)
    _, e_gb = H.minimize_mol(mol, "gbneck_tip3p", model_path={"model": sd})
    assert np.all(np.isfinite(e_vac)) and np.all(np.isfinite(e_gb)) and np.all(e_gb < e_vac)   # solvation lowers the energy


@pytest.mark.gpu
def test_calculate_entropy_entry_point():
    sd = _sd()
    mol = synth.synth_mol_object(10, 3, seed=11, sigma=0.02)
    S, G_ = H.calculate_entropy(mol, "tip3p", model_path={"model": sd})
    assert S.shape == (3,) and G_.shape == (3,) and np.all(np.isfinite(S)) and np.all(S > 0)
    only_s = H.calculate_entropy(synth.synth_mol_object(10, 3, seed=11, sigma=0.02), "tip3p", model_path={"model": sd},
                                 return_free_energy=False)
    assert np.allclose(only_s, S, rtol=1e-3)


@pytest.mark.gpu
def test_run_simulation_writes_trajectory_and_log(tmp_path):
    sd = _sd()
    mol = synth.synth_mol_object(13, 2, seed=13)
    sim = Multi_simulator(work_dir=str(tmp_path) + "/", mol=mol, state_dict=sd, solvent_model=[0, 3],
                          solvent_dielectric=[78.5, 47.24], num_rep=2, random_number_seed=4)
    sim.setup_replicates(only_check_first=True)
    prefix = sim.run_simulation(n_steps=40, n_interval=10, minimize=True)
    from gnnimplicitsolvent_b200 import trajectory
    xyz, meta = trajectory.load_trajectory(prefix)
    assert xyz.shape[0] == 4 and np.all(np.isfinite(np.asarray(xyz)))
    log = open(sim.log_loc).read().strip().splitlines()
    assert len(log) == 5 and log[0].startswith('#"Step"')
    assert sim.hdf5_loc == prefix + "_output.h5" and sim.log_loc == prefix + "_log.txt"
    # the HDF5 container the reference's analysis reads (analysis.py:2032-2037): one chain per replica
    per_rep = trajectory.load_parallel_traj(sim.hdf5_loc, 2)
    np.testing.assert_array_equal(per_rep[1], np.asarray(xyz)[:, 13:26])


@pytest.mark.gpu
def test_create_gnn_model_entry_point():
    sd = _sd()
    mol = synth.synth_mol_object(21, 1, seed=5, with_forcefield=False)
    model = G.create_gnn_model(mol, solvent="DMSO", model_dict=sd, device="cuda")
    pos = torch.from_numpy(mol.conformers_nm[0].astype(np.float32)).cuda().requires_grad_(True)
    e = model(pos, torch.tensor(-1), torch.tensor(78.5))
    e.backward()
    from gnnimplicitsolvent_b200.engine import Engine
    eng = Engine(mol.parameters, sd)
    eng.set_replicas(1, [3], [47.24])
    e2, f2 = eng.energy_force(pos.detach().reshape(1, 21, 3))
    assert abs(float(e) - float(e2[0])) < 1e-4 * abs(float(e2[0])) + 1e-4
    assert torch.allclose(-pos.grad, f2[0], rtol=1e-4, atol=1e-3)


def test_helper_plumbing_without_gpu():
    """Solvent lists (interleaving of helper_functions.py:738-743), trajectory assembly and the write-back into the molecule."""
    sm, sd = H._solvent_lists("DMSO", H.SOLVENT_DICT, 3)
    assert sm == [3, 3, 3] and sd == [47.24] * 3
    sm, sd = H._solvent_lists(["tip3p", "DMSO"], H.SOLVENT_DICT, 4)
    assert sm == [0, 3, 0, 3] and sd == [78.5, 47.24, 78.5, 47.24]
    mol = synth.synth_mol_object(7, 5, seed=1)
    blocks = [np.full((2 * 7, 3), 1.0), np.full((2 * 7, 3), 2.0), np.full((1 * 7, 3), 3.0)]     # strides=2 + one left over
    traj = H.get_traj_from_positions(mol, blocks, num_iters=2, num_confs=2)
    assert traj.n_frames == 5 and traj.n_atoms == 7
    assert np.all(traj.xyz[0] == 1.0) and np.all(traj.xyz[3] == 2.0) and np.all(traj.xyz[4] == 3.0)
    H.set_positions_in_mol(mol, [np.concatenate(blocks, axis=0)])
    assert np.allclose(mol.conformers_nm[2], 2.0) and np.allclose(mol.GetConformer(4).GetPositions(), 30.0)
    assert abs(H.get_boltzmann_mean(np.zeros(4))) < 1e-12


@pytest.mark.gpu
def test_entropy_reminimises_conformers_with_imaginary_modes():
    """calculate_entropy_from_simulation on conformers that are NOT at a minimum: their lowest mode is imaginary, and
    the reference's minimization_cycles loop (helper_functions.py:1986-2015) re-minimises exactly those until the
    modes are real; conformers already at a minimum are left alone."""
    sd = _sd()
    mol = synth.synth_mol_object(10, 4, seed=11, sigma=0.02)
    sim = H.get_gnn_sim(mol, "tip3p", {"model": sd}, H.SOLVENT_DICT, constraints=None, num_confs=4)
    start = mol.conformers_nm.astype(np.float32)
    from gnnimplicitsolvent_b200 import entropy as E
    _, nus0 = E.conformer_entropies(sim._ref_system._engine, start, mol.masses, 0, 78.5)
    assert any(nu[0] < 0 for nu in nus0)                       # displaced start points: imaginary modes
    S, nus, pos = H.calculate_entropy_from_simulation(sim._ref_system, start)
    assert np.all(np.isfinite(S)) and S.shape == (4,)
    assert sum(nu[0] > 0 for nu in nus) >= 3                   # re-minimised: real lowest modes (a ledge may keep one)
    assert not np.allclose(pos, start)
    S2, nus2, pos2 = H.calculate_entropy_from_simulation(sim._ref_system, pos, use_minimisation_cycles=False)
    assert np.allclose(S2, S, rtol=1e-4)                       # without the loop the (already minimised) input is evaluated as it is
