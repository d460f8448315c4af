"""The reference's named entry points (Simulation/helper_functions.py) on the resident batched engine:

    minimize_mol            :931-1129   conformers of a molecule minimised in a solvent, energies per conformer
    get_gnn_sim / create_gnn_sim / create_vac_sim   :698-790, :1132-1184
    set_positions_for_simulation / run_minimisation / get_minimised_positions   :600-632
    get_energies            :901-918
    calculate_entropy       :1688-1745 (+ minimization_cycles :1986-2015)
    create_gnn_model        :826-898

Argument lists, defaults, return shapes and the failure semantics (a failed batch is re-minimised conformer by
conformer, a conformer that fails again comes back as zeros, :1021-1054) are the reference's.  What differs is what
a `mol` is: RDKit / OpenFF / OpenMM are not in this image, so `mol` is a `Simulator.Molecule` (element symbols,
conformers, the [N,7] GB-Neck2 table, optional exported force-field terms) instead of an RDKit molecule;
`partial_charges`, `forcefield` and `cache` are accepted and ignored when the molecule already carries its
parameters.  Energies are kJ/mol, coordinates nm (RDKit-style conformer access in Angstrom)."""
from __future__ import annotations

import os
import tempfile
import warnings
from typing import Optional

import numpy as np
import torch
import yaml

from . import entropy as _entropy
from . import trajectory as _trajectory
from .GNN_Models import GNN3_Multisolvent_embedding_run_multiple
from .GNN_Models import create_gnn_model as _create_model_from_parameters
from .Simulator import HBonds, Molecule, Multi_simulator, Simulator  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
SOLVENT_DICT = yaml.safe_load(open(os.path.join(_HERE, "data", "solvents.yml")))["solvent_mapping_dict"]


def _find_model_path():
    """The shipped checkpoint: trained_models/GNN.pt of a reference checkout when one is reachable, otherwise the
    same state dict as stored next to the golden vectors (tests/golden/gnn_weights.npz)."""
    root = os.path.dirname(_HERE)
    for p in (os.path.join(root, "baseline", "_ref", "MachineLearning", "trained_models", "GNN.pt"),
              "/root/reference/MachineLearning/trained_models/GNN.pt",
              os.path.join(_HERE, "data", "gnn_weights.npz"),
              os.path.join(root, "tests", "golden", "gnn_weights.npz")):
        if os.path.exists(p):
            return p
    return os.path.join(root, "tests", "golden", "gnn_weights.npz")


MODEL_PATH = _find_model_path()


def load_model_dict(model_path=MODEL_PATH):
    """State dict of a checkpoint: a `.pt` file with a "model" entry (the reference's format) or an `.npz` of it."""
    if isinstance(model_path, dict):
        return model_path.get("model", model_path)
    if str(model_path).endswith(".npz"):
        z = np.load(model_path)
        return {k: torch.from_numpy(z[k]) for k in z.files}
    sd = torch.load(model_path, map_location="cpu", weights_only=True)
    return sd["model"] if "model" in sd else sd


class Trajectory:
    """The two mdtraj.Trajectory attributes the entry points use: `xyz` [frames, atoms, 3] (nm) and `n_frames`."""

    def __init__(self, xyz, elements=None):
        self.xyz = np.asarray(xyz, dtype=np.float32)
        self.elements = elements

    @property
    def n_frames(self):
        return int(self.xyz.shape[0])

    @property
    def n_atoms(self):
        return int(self.xyz.shape[1])

    def __getitem__(self, i):
        x = self.xyz[i]
        return Trajectory(x if x.ndim == 3 else x[None], self.elements)


def get_boltzmann_mean(energies):
    return -np.log(np.mean(np.exp(-energies)))


def kjmol_to_prop(energies):
    kT = 2.479
    return np.exp(-energies / kT) / np.sum(np.exp(-energies / kT))


# ---------------------------------------------------------------------------------------------- simulators
def _solvent_lists(solvent, solvent_dict, num_confs):
    if isinstance(solvent, str):
        return ([solvent_dict[solvent]["solvent_id"]] * num_confs, [solvent_dict[solvent]["dielectric"]] * num_confs)
    sm, sd = [], []
    for _ in range(num_confs // len(solvent)):       # interleaved as helper_functions.py:738-743
        for s in solvent:
            sm.append(solvent_dict[s]["solvent_id"])
            sd.append(solvent_dict[s]["dielectric"])
    return sm, sd


def create_gnn_sim(smiles, cache=None, num_confs=64, workdir=None, run_name="", save_name=None, rdkit_mol=None,
                   solvent_dict=None, solvent="tip3p", num_solvents=42, model_dict=None, solvent_model=None,
                   solvent_dielectric=None, partial_charges=None, forcefield="openff-2.0.0", constraints=HBonds,
                   device=0):
    """Multi_simulator of `num_confs` replicas of `rdkit_mol` (a Molecule) with the GNN model (:698-823)."""
    workdir = workdir if workdir is not None else tempfile.mkdtemp() + "/"
    solvent_dict = solvent_dict if solvent_dict is not None else SOLVENT_DICT
    if solvent_model is None:
        solvent_model, solvent_dielectric = _solvent_lists(solvent, solvent_dict, num_confs)
    if isinstance(model_dict, str):
        model_dict = load_model_dict(model_dict)
    gnn_sim = Multi_simulator(work_dir=workdir, pdb_id=str(smiles) + "_in_v", mol=rdkit_mol, state_dict=model_dict,
                              solvent_model=solvent_model, solvent_dielectric=solvent_dielectric, num_rep=num_confs,
                              run_name=run_name, save_name=save_name, cache=cache, partial_charges=partial_charges,
                              openff_forcefield=forcefield, constraints=constraints, device=device)
    gnn_sim.setup_replicates(only_check_first=True)
    return gnn_sim


def create_vac_sim(smiles, cache=None, num_confs=64, save_name=None, mol=None, partial_charges=None,
                   forcefield="openff-2.0.0", constraints=HBonds, device=0):
    """Vacuum force field only (solvent == "vac", :1159-1169); needs the molecule's exported force-field terms."""
    sim = Multi_simulator(pdb_id=str(smiles) + "_in_v", mol=mol, num_rep=num_confs, save_name=save_name, cache=cache,
                          partial_charges=partial_charges, openff_forcefield=forcefield, constraints=constraints,
                          device=device, model="vacuum")
    sim.setup_replicates(only_check_first=True)
    return sim


def create_gbneck_sim(smiles, solvent, solvent_dict, num_confs=64, save_name=None, mol=None, constraints=HBonds, device=0):
    """Plain GB-Neck2 (the `"gbneck" in solvent` branch of minimize_mol, :969-984): solvent "gbneck_<name>"."""
    name = solvent.replace("gbneck_", "").replace("gbneck", "") or "tip3p"
    die = solvent_dict[name]["dielectric"] if name in solvent_dict else 78.5
    sim = Multi_simulator(pdb_id=str(smiles) + "_in_v", mol=mol, solvent_model=[0] * num_confs,
                          solvent_dielectric=[die] * num_confs, num_rep=num_confs, save_name=save_name,
                          constraints=constraints, device=device, model="gbneck")
    sim.setup_replicates(only_check_first=True)
    return sim


def get_gnn_sim(mol, solvent, model_path, solvent_dict, cache=None, save_name=None, partial_charges=None,
                forcefield="openff-2.0.0", constraints=HBonds, num_confs=1, device=0):
    """:1132-1184"""
    smiles = getattr(mol, "smiles", "MOL")
    if solvent == "vac":
        return create_vac_sim(smiles, cache, num_confs, save_name, mol=mol, partial_charges=partial_charges,
                              forcefield=forcefield, constraints=constraints, device=device)
    if isinstance(solvent, str) and "gbneck" in solvent:
        return create_gbneck_sim(smiles, solvent, solvent_dict, num_confs, save_name, mol=mol, constraints=constraints,
                                 device=device)
    return create_gnn_sim(smiles, cache=cache, num_confs=num_confs, workdir=tempfile.mkdtemp() + "/",
                          run_name="testminimise", save_name=save_name, rdkit_mol=mol, solvent=solvent,
                          solvent_dict=solvent_dict, model_dict=load_model_dict(model_path),
                          partial_charges=partial_charges, forcefield=forcefield, constraints=constraints, device=device)


def set_positions_for_simulation(sim, mol, num_confs=64, iteration=0, offset=0):
    """Transfer num_confs conformers from mol to sim, starting from num_confs*iteration + offset (:600-610)."""
    positions = []
    for i in range((iteration * num_confs) + offset, ((iteration + 1) * num_confs + offset)):
        positions.append(mol.GetConformer(i).GetPositions() / 10)
    positions = np.array(positions)
    sim.set_positions(positions.reshape(-1, positions.shape[-1]))
    return sim


def run_minimisation(sim, tolerance=1e-4, max_iterations=0):
    """:613-624 -- status 0 when every replica of the batch ended with status 0, else 1 (the reference's bare
    `except` turns any failure of the OpenMM call into 1); the per-replica status stays in sim.last_minimize_status."""
    status = 0
    try:
        sim._simulation.minimizeEnergy(tolerance=tolerance, maxIterations=max_iterations)
        if sim.last_minimize_status is not None and np.any(sim.last_minimize_status == 2):
            status = 1     # a replica went non-finite: what raises inside OpenMM
    except Exception:      # noqa: BLE001  (mirrors the reference)
        status = 1
    return sim, status


def get_minimised_positions(sim):
    return sim.pos


def get_traj_from_positions(mol, positions, num_iters=1, num_confs=64):
    n_atoms = mol.GetNumAtoms()
    xyz = np.zeros((mol.GetNumConformers(), n_atoms, 3), dtype=np.float32)
    k = 0
    for block in positions:
        b = np.asarray(block).reshape(-1, n_atoms, 3)
        xyz[k:k + len(b)] = b
        k += len(b)
    return Trajectory(xyz, getattr(mol, "elements", None))


def set_positions_in_mol(mol, positions):
    """Use an array of shape (1, n_atoms*n_frames, 3) to update n_frames conformers (:652-666)."""
    n_confs, n_atoms = mol.GetNumConformers(), mol.GetNumAtoms()
    max_n_confs = len(positions[0]) // n_atoms
    if max_n_confs < n_confs:
        print(f"Warning: too few positions in set_positions_in_mol; the last {n_confs - max_n_confs} conformers will not be updated.")
    for i in range(min(n_confs, max_n_confs)):
        conf = mol.GetConformer(i)
        for j in range(n_atoms):
            conf.SetAtomPosition(j, positions[0][i * n_atoms + j] * 10)
    return mol


def get_energies(sim, traj, additional_parameters=[]):
    """Potential energy of every frame on the single-copy reference system (:901-918): here all frames are replicas
    of batched calls of that system's engine; a frame that does not evaluate comes back as NaN."""
    ref = sim._ref_system
    if len(additional_parameters) != 0:
        out = []
        for i in range(traj.n_frames):           # per-frame global parameters: solvent_model / solvent_dielectric
            p = additional_parameters[i]
            out.append(_trajectory.get_energies(ref._engine, traj.xyz[i:i + 1],
                                                int(p.get("solvent_model", ref._solvent_model[0])),
                                                float(p.get("solvent_dielectric", ref._solvent_dielectric[0])),
                                                **ref._flags())[0])
        e = np.array(out)
    else:
        e = _trajectory.get_energies(ref._engine, traj.xyz, ref._solvent_model[0], ref._solvent_dielectric[0], **ref._flags())
    ref._engine.set_replicas(1, [ref._solvent_model[0]], [ref._solvent_dielectric[0]])
    bad = ~np.isfinite(e)
    if bad.any():
        warnings.warn(f"Energy calculation failed for frames {np.nonzero(bad)[0].tolist()}")
        e[bad] = np.nan
    return e


# ---------------------------------------------------------------------------------------------- minimize_mol
def _minimise_block(gnn_sim, mol, num_confs, iteration, offset, tolerance, max_iterations, make_single):
    """One batch of `num_confs` conformers; failed replicas are redone alone (reference: :1021-1054)."""
    n_atoms = mol.GetNumAtoms()
    gnn_sim = set_positions_for_simulation(gnn_sim, mol, num_confs=num_confs, iteration=iteration, offset=offset)
    gnn_sim, status = run_minimisation(gnn_sim, tolerance=tolerance, max_iterations=max_iterations)
    pos = get_minimised_positions(gnn_sim).reshape(num_confs, n_atoms, 3).copy()
    rep_status = gnn_sim.last_minimize_status if gnn_sim.last_minimize_status is not None else np.ones(num_confs, np.int32)
    failed = np.nonzero(rep_status == 2)[0] if status != 0 else []
    if len(failed):
        print(f"Batch Optimization failed for iteration {iteration}, starting individual optimization")
        single = make_single()
        for j in failed:
            single = set_positions_for_simulation(single, mol, num_confs=1, iteration=0,
                                                  offset=iteration * num_confs + offset + int(j))
            single, st = run_minimisation(single, tolerance=tolerance, max_iterations=max_iterations)
            if st == 0:
                pos[j] = get_minimised_positions(single).reshape(n_atoms, 3)
            else:
                print(f"Individual optimization failed for confomer {iteration * num_confs + offset + int(j)}")
                pos[j] = 0.0
        del single
    return pos.reshape(-1, 3), rep_status


def minimize_mol(mol, solvent, model_path=MODEL_PATH, solvent_dict=SOLVENT_DICT, return_traj=False, tolerance=1e-4,
                 max_iterations=0, gnn_sim=None, return_gnn_sim=False, strides=1, cache=None, save_name=None,
                 partial_charges=None, forcefield="openff-2.0.0", constraints=HBonds, device=0):
    """Minimize each conformer of a molecule (:931-1129).

    mol: Molecule with at least one conformer; solvent: key of solvent_dict, "vac", or "gbneck_<key>";
    strides: number of groups the conformers are minimised in (all conformers of a group are resident at once).
    Returns (mol, energies) | (mol, traj, energies) | (mol, traj, energies, gnn_sim) like the reference; energies are
    the per-conformer potential energies of the single-copy reference system at the minimised geometries."""
    num_rep = mol.GetNumConformers()
    num_iters = strides
    num_confs = num_rep // strides
    kw = dict(mol=mol, solvent=solvent, model_path=model_path, solvent_dict=solvent_dict, cache=cache, save_name=save_name,
              partial_charges=partial_charges, forcefield=forcefield, constraints=constraints, device=device)
    if gnn_sim is None:
        gnn_sim = get_gnn_sim(num_confs=num_confs, **kw)
    positions, statuses = [], []
    for i in range(num_iters):
        p, s = _minimise_block(gnn_sim, mol, num_confs, i, 0, tolerance, max_iterations, lambda: get_gnn_sim(num_confs=1, **kw))
        positions.append(p)
        statuses.append(s)
    n_missing = num_rep - num_iters * num_confs
    if n_missing:
        print(f"Individually optimizing the last {n_missing} conformers")
        extra = get_gnn_sim(num_confs=n_missing, **kw)
        p, s = _minimise_block(extra, mol, n_missing, 0, num_iters * num_confs, tolerance, max_iterations,
                               lambda: get_gnn_sim(num_confs=1, **kw))
        positions.append(p)
        statuses.append(s)
        del extra
    gnn_traj = get_traj_from_positions(mol, positions, num_iters=num_iters, num_confs=num_confs)
    gnn_energies = get_energies(gnn_sim, gnn_traj)
    reshaped_positions = np.concatenate(positions, axis=0)
    mol = set_positions_in_mol(mol, [reshaped_positions])
    gnn_sim.minimize_status = np.concatenate(statuses)       # per-conformer status of this call (0 / 1 / 2)
    if return_traj and return_gnn_sim:
        return mol, gnn_traj, gnn_energies, gnn_sim
    elif return_traj:
        return mol, gnn_traj, gnn_energies
    elif return_gnn_sim:
        return mol, gnn_energies, gnn_sim
    return mol, gnn_energies


# ---------------------------------------------------------------------------------------------- entropy
def calculate_entropy_from_simulation(ref_sim: Simulator, positions_nm, delta=0.0001, use_minimisation_cycles=True,
                                      max_cycles=100):
    """T*S of every conformer in `positions_nm` [C,N,3] on ref_sim's Hamiltonian (:1748-1763), with the reference's
    re-minimisation loop for conformers whose lowest mode is imaginary (minimization_cycles, :1986-2015): minimise
    again with the tighter tolerance, recompute the modes when the mean force moved by more than 0.1, stop when the
    lowest mode is real or after 100 cycles.  All conformers (and all their 6N displaced copies) are batched."""
    eng, masses = ref_sim._engine, ref_sim._mol.masses
    sid, die = ref_sim._solvent_model[0], ref_sim._solvent_dielectric[0]
    pos = np.array(positions_nm, dtype=np.float32)
    flags = ref_sim._flags()
    S, nus = _entropy.conformer_entropies(eng, pos, masses, sid, die, delta, **flags)
    if use_minimisation_cycles:
        imag = np.array([nu[0] < 0 for nu in nus])          # imaginary modes are reported as negative wavenumbers
        old_mf = None
        for _ in range(max_cycles):
            if not imag.any():
                break
            idx = np.nonzero(imag)[0]
            eng.set_replicas(len(idx), [sid] * len(idx), [die] * len(idx))
            p, _, _, _ = eng.minimize(torch.from_numpy(pos[idx]).to(eng.device), grad_tol=0.1, max_iter=2000)
            _, f = eng.energy_force(p, **flags)
            mf = f.norm(dim=2).mean(dim=1).cpu().numpy()
            pos[idx] = p.cpu().numpy()
            moved = np.ones(len(idx), bool) if old_mf is None or len(old_mf) != len(mf) else np.abs(mf - old_mf) >= 0.1
            old_mf = mf
            if not moved.any():
                continue
            S2, nus2 = _entropy.conformer_entropies(eng, pos[idx[moved]], masses, sid, die, delta, **flags)
            for k, j in enumerate(idx[moved]):
                S[j], nus[j] = S2[k], nus2[k]
                imag[j] = nus2[k][0] < 0
        if imag.any():
            warnings.warn("No real frequencies found after 100 minimisation cycles.")
    eng.set_replicas(1, [sid], [die])
    return S, nus, pos


def calculate_entropy(mol, solvent, model_path=MODEL_PATH, solvent_dict=SOLVENT_DICT, tolerance=1e-4, max_iterations=0,
                      gnn_sim=None, strides=1, cache=None, save_name=None, partial_charges=None,
                      forcefield="openff-2.0.0", return_free_energy=True, return_traj=False, device=0):
    """:1688-1745 -- minimise every conformer without constraints, then T*S per conformer (quasi-RRHO vibrations +
    rigid rotor) from the batched Hessians.  Returns (entropies, energies - entropies), (traj, energies - entropies)
    or entropies alone, as the reference."""
    mol, gnn_traj, gnn_energies, gnn_sim = minimize_mol(
        mol, solvent, model_path, solvent_dict, return_traj=True, tolerance=tolerance, max_iterations=max_iterations,
        gnn_sim=gnn_sim, return_gnn_sim=True, strides=strides, cache=cache, partial_charges=partial_charges,
        save_name=save_name, forcefield=forcefield, constraints=None, device=device)
    try:
        entropies, _, _ = calculate_entropy_from_simulation(gnn_sim._ref_system, gnn_traj.xyz)
    except Exception as exc:  # noqa: BLE001
        warnings.warn(f"Entropy calculation failed: {exc}")
        entropies = np.zeros(mol.GetNumConformers())
    entropies = np.asarray(entropies, dtype=np.float64)
    if return_free_energy:
        if return_traj:
            return gnn_traj, gnn_energies - entropies
        return entropies, gnn_energies - entropies
    return entropies


# ---------------------------------------------------------------------------------------------- model
def create_gnn_model(mol, model_class=GNN3_Multisolvent_embedding_run_multiple, cache=None, solvent_dict=SOLVENT_DICT,
                     solvent="tip3p", num_solvents=42, model_dict=None, partial_charges=None,
                     forcefield="openff-2.0.0", device="cuda", jit=False):
    """Run model for ONE copy of `mol` in `solvent` with the checkpoint loaded (:826-898).  `mol` is a Molecule or a
    bare [N,7] parameter table; `model_dict` a state dict, a path, or None for the shipped checkpoint."""
    if model_dict is None or isinstance(model_dict, str):
        model_dict = load_model_dict(model_dict if model_dict is not None else MODEL_PATH)
    parameters = mol.parameters if hasattr(mol, "parameters") else np.asarray(mol)
    sm, sd = _solvent_lists(solvent, solvent_dict, 1 if isinstance(solvent, str) else len(solvent))
    return _create_model_from_parameters(parameters, model_dict, num_reps=1, solvent_models=[sm[0]],
                                         solvent_dielectric=[sd[0]], cls=model_class, device=device, jit=jit)
