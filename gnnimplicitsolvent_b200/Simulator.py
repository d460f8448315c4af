"""Simulator / Multi_simulator -- the reference's simulation front end (Simulation/Simulator.py:70-946, 1522-1711)
on top of the resident batched engine.

The reference builds an OpenMM `Simulation` around a TorchScript force and lets OpenMM's minimiser / integrator call
the model once per step.  Here the whole replicated system (vacuum force field + GB-Neck2 + GNN) lives behind one
libgnnis handle and the minimiser / Langevin integrator are the batched CUDA loops of csrc/dynamics.cu, so the
classes keep the reference's method names and semantics while nothing crosses the Python boundary inside a step:

    Simulator          one copy of the molecule  (`_ref_system` of a Multi_simulator)
    Multi_simulator    `num_rep` replicas, `setup_replicates()` with the reference's force self-check
    set_positions / calculate_energy / calculate_forces / run_simulation(n_steps, n_interval, minimize) / pos /
    hdf5_loc / log_loc / `_simulation.minimizeEnergy(tolerance, maxIterations)` / `_simulation.step(n)`

RDKit, OpenFF and OpenMM are absent from this image, so a molecule is the light `Molecule` record below: element
symbols, conformer coordinates, the [N,7] GB-Neck2 parameter table and, optionally, the vacuum force-field terms
exported once from an OpenMM System (`Engine.set_forcefield` layout).  Units follow the reference: nm, kJ/mol, ps.
"""
from __future__ import annotations

import os
import tempfile
from typing import Optional, Sequence

import numpy as np
import torch

from . import trajectory
from .engine import Engine

ELEMENT_MASS = {"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "F": 18.998, "P": 30.974, "S": 32.06, "Cl": 35.45,
                "Br": 79.904, "I": 126.904}


class _Conformer:
    def __init__(self, mol, i):
        self._mol, self._i = mol, i

    def GetPositions(self):          # RDKit convention: Angstrom
        return self._mol.conformers_nm[self._i] * 10.0

    def SetAtomPosition(self, j, xyz):
        self._mol.conformers_nm[self._i][j] = np.asarray([xyz[0], xyz[1], xyz[2]], dtype=np.float64) / 10.0


class Molecule:
    """What the entry points need from an RDKit mol + OpenFF parametrisation, as plain arrays.

    elements      [N] element symbols
    conformers_nm [C,N,3] conformer coordinates in nm (GetConformer(i).GetPositions() returns Angstrom like RDKit)
    parameters    [N,7] GB-Neck2 table [q, rho, s, alpha, beta, gamma, radindex] (GNN_Trainer.py:829-869)
    forcefield    optional dict of vacuum terms (Engine.set_forcefield); constraints are derived from its X-H bonds
    """

    def __init__(self, elements: Sequence[str], conformers_nm, parameters, forcefield: Optional[dict] = None,
                 masses: Optional[Sequence[float]] = None, smiles: str = "MOL"):
        self.elements = [str(e) for e in elements]
        self.conformers_nm = np.array(conformers_nm, dtype=np.float64).reshape(-1, len(self.elements), 3)
        self.parameters = np.asarray(parameters, dtype=np.float64)
        self.forcefield = forcefield
        self.masses = np.asarray(masses if masses is not None else
                                 [ELEMENT_MASS.get(e.rstrip("0123456789").replace("n", ""), 12.011) for e in self.elements],
                                 dtype=np.float64)
        self.smiles = smiles

    def GetNumConformers(self):
        return int(self.conformers_nm.shape[0])

    def GetNumAtoms(self):
        return len(self.elements)

    def GetConformer(self, i):
        return _Conformer(self, int(i))

    def hbond_constraints(self):
        """(idx [K,2], d [K]) of all bonds to hydrogen at their force-field length (OpenMM `constraints=HBonds`)."""
        if self.forcefield is None or "bonds" not in self.forcefield:
            return np.zeros((0, 2), np.int32), np.zeros((0,), np.float32)
        idx = np.asarray(self.forcefield["bonds"]["idx"]).reshape(-1, 2)
        r0 = np.asarray(self.forcefield["bonds"]["r0"]).reshape(-1)
        sel = [k for k, (i, j) in enumerate(idx) if self.elements[i].startswith("H") or self.elements[j].startswith("H")]
        return idx[sel].astype(np.int32), r0[sel].astype(np.float32)


HBonds = "HBonds"   # stands in for openmm.app.HBonds in argument lists


class _SimulationShim:
    """`sim._simulation` of the reference: the two OpenMM calls the helper functions make on it."""

    def __init__(self, owner):
        self._o = owner
        self.reporters = []

    def minimizeEnergy(self, tolerance=1e-4, maxIterations=0):
        self._o.minimize(tolerance=tolerance, max_iterations=maxIterations)

    def step(self, n_steps):
        self._o._advance(int(n_steps))


class Simulator:
    """One resident system of `num_rep` copies of a molecule (1 for the plain Simulator)."""

    def __init__(self, work_dir: str = None, name: str = "", pdb_id: str = "", mol: Molecule = None, state_dict=None,
                 solvent_model: Sequence[int] = (0,), solvent_dielectric: Sequence[float] = (78.5,), num_rep: int = 1,
                 run_name: str = "", save_name: Optional[str] = None, random_number_seed: int = 0, constraints=HBonds,
                 device=0, temperature: float = 300.0, friction: float = 1.0, dt: float = 0.002, model: str = "gnn"):
        if mol is None:
            raise ValueError("Simulator needs a Molecule (RDKit / OpenFF are not available to build one from SMILES)")
        work_dir = work_dir if work_dir is not None else tempfile.mkdtemp() + "/"
        self._work_folder = work_dir + "Simulation/simulation/" + run_name + "/"
        os.makedirs(self._work_folder, exist_ok=True)
        self._work_dir, self._run_name = work_dir, run_name
        self._name, self._pdb_id = name, pdb_id or (mol.smiles + "_in_v")
        self._save_name = save_name if save_name is not None else self._pdb_id
        self._iteration = 0
        self._random_number_seed = int(random_number_seed)
        self._mol, self._state_dict, self._device = mol, state_dict, device
        self._num_rep = int(num_rep)
        self._constraints = constraints
        if model not in ("gnn", "gbneck", "vacuum"):
            raise ValueError("model must be 'gnn', 'gbneck' or 'vacuum'")
        if model == "gnn" and state_dict is None:
            raise ValueError("the GNN model needs a state dict")
        self._model = model          # Hamiltonian on top of the vacuum terms: GNN implicit solvent, plain GB-Neck2, nothing
        vacuum = model == "vacuum"
        self._vacuum = vacuum
        self._temperature, self._friction, self._dt = temperature, friction, dt
        self._forcefield_name = {"gnn": "GNN3_vap", "gbneck": "GBNeck2_vap", "vacuum": "vacuum"}[model]
        self._engine = Engine(mol.parameters, state_dict if model == "gnn" else None, device=device)
        if mol.forcefield is not None:
            self._engine.set_forcefield(mol.forcefield)
        self._n_constraints = 0
        if constraints is not None and mol.forcefield is not None:
            idx, d = mol.hbond_constraints()
            if len(idx):
                self._engine.set_constraints(idx, d)
                self._n_constraints = len(idx)
        if vacuum and mol.forcefield is None:
            raise ValueError("a vacuum simulator needs force-field terms")
        sm = list(solvent_model) * (self._num_rep // max(1, len(solvent_model))) if len(solvent_model) != self._num_rep else list(solvent_model)
        sd = list(solvent_dielectric) * (self._num_rep // max(1, len(solvent_dielectric))) if len(solvent_dielectric) != self._num_rep else list(solvent_dielectric)
        if len(sm) != self._num_rep or len(sd) != self._num_rep:
            raise ValueError("need one solvent model / dielectric per replica (or a list that divides num_rep)")
        self._solvent_model, self._solvent_dielectric = sm, sd
        self._engine.set_replicas(self._num_rep, sm, sd)
        N = mol.GetNumAtoms()
        self._positions = torch.zeros(self._num_rep, N, 3, dtype=torch.float32, device=self._engine.device)
        self._velocities = torch.zeros_like(self._positions)
        c0 = torch.as_tensor(mol.conformers_nm[0], dtype=torch.float32)
        self._positions.copy_(c0.unsqueeze(0).expand(self._num_rep, N, 3))
        self._steps_done = 0
        self._writer = None
        self._simulation = _SimulationShim(self)
        self.last_minimize_status = None

    # ------------------------------------------------------------------ reference surface
    def __str__(self):
        return self._name

    @property
    def pos(self):
        """Positions [num_rep * N, 3] in nm (the reference returns the context's positions the same way)."""
        return self._positions.reshape(-1, 3).cpu().numpy().astype(np.float64)

    def set_positions(self, positions):
        """positions [num_rep * N, 3] (or [num_rep, N, 3]) in nm (Simulator.py:828-831)."""
        p = torch.as_tensor(np.asarray(positions), dtype=torch.float32).reshape(self._positions.shape)
        self._positions.copy_(p)

    def _flags(self):
        return {"vacuum": dict(vacuum_only=True), "gbneck": dict(gb_only=True), "gnn": {}}[self._model]

    def calculate_energy(self) -> float:
        """Total potential energy of the system in kJ/mol (Simulator.py:800-804: the sum over all replicas)."""
        return float(self.calculate_energies().sum())

    def calculate_energies(self) -> np.ndarray:
        """Per-replica potential energies [num_rep] (the engine returns them natively)."""
        e, _ = self._engine.energy_force(self._positions, forces=False, **self._flags())
        return e.double().cpu().numpy()

    def calculate_forces(self, adapt_values=False, use_NN=False) -> np.ndarray:
        """Forces as the reference returns them: array [3, num_rep * N] (Simulator.py:768-798), kJ mol^-1 nm^-1."""
        _, f = self._engine.energy_force(self._positions, **self._flags())
        return f.reshape(-1, 3).t().contiguous().double().cpu().numpy()

    def minimize(self, tolerance=1e-4, max_iterations=0):
        """`simulation.minimizeEnergy(tolerance, maxIterations)` (helper_functions.py:614-624).  OpenMM's tolerance
        is a gradient-RMS bound in kJ mol^-1 nm^-1 that the reference sets far below the fp32 noise floor (1e-4):
        every replica then runs until its line search stalls, which is what the batched L-BFGS does for
        grad_tol <= 1.  Returns per-replica status (0 gradient RMS below grad_tol, 1 iteration limit, 2 non-finite,
        3 line search exhausted at fp32 resolution above grad_tol -- the reference's normal way to stop)."""
        grad_tol = max(float(tolerance), 1.0)
        max_iter = int(max_iterations) if max_iterations and max_iterations > 0 else 2000
        p, e, status, rounds = self._engine.minimize(self._positions, grad_tol=grad_tol, max_iter=max_iter, **self._flags())
        self._positions.copy_(p)
        self.last_minimize_status = status.cpu().numpy()
        self.last_minimize_rounds = rounds
        self.last_minimize_energies = e.double().cpu().numpy()
        return self.last_minimize_status

    def _file_stub(self):
        return (self._save_name + "_" + self._forcefield_name + "_" + str(self._iteration) + "_" +
                str(self._random_number_seed))

    @property
    def hdf5_loc(self):
        """Trajectory location (Simulator.py:405-417): the MDTraj-convention HDF5 file run_simulation leaves behind.  During
        the run the frames go to `<stub>_output.coordinates.npy` + `<stub>_output.meta.npz` (trajectory.TrajectoryWriter,
        memory-mapped); trajectory.export_mdtraj_h5 converts them to this path when the run ends."""
        return self._work_folder + self._file_stub() + "_output.h5"

    @property
    def log_loc(self):
        return self._work_folder + self._file_stub() + "_log.txt"

    def _advance(self, n_steps: int):
        self._engine.langevin(self._positions, self._velocities, self._mol.masses, n_steps,
                              temperature=self._temperature, friction=self._friction, dt=self._dt,
                              seed=self._random_number_seed, first_step=self._steps_done, **self._flags())
        self._steps_done += n_steps

    def run_simulation(self, n_steps: int = 10000, n_interval: int = 1000, minimize: bool = True, workfolder=None):
        """Simulator.run_simulation (Simulator.py:341-402): optional minimisation on the first call, LangevinMiddle
        (300 K, 1/ps, 2 fs) with a frame and a StateDataReporter-style log line every n_interval steps."""
        if workfolder is None:
            workfolder = self._work_folder
        elif workfolder == "TMPDIR":
            workfolder = os.environ["TMPDIR"] + "/"
        os.makedirs(workfolder, exist_ok=True)
        prefix = workfolder + self._file_stub()     # TrajectoryWriter: <prefix>_output.coordinates.npy / .meta.npz, <prefix>_log.txt
        do_min = bool(minimize and self._iteration == 0)
        if do_min:
            self.minimize()
        trajectory.run_simulation(self._engine, self._positions, self._velocities, self._mol.masses, n_steps=n_steps,
                                  n_interval=n_interval, minimize=False, prefix=prefix, temperature=self._temperature,
                                  friction=self._friction, dt=self._dt, seed=self._random_number_seed,
                                  elements=self._mol.elements, n_constraints=self._n_constraints,
                                  first_step=self._steps_done, eval_flags=self._flags())
        self._steps_done += (n_steps // n_interval) * n_interval
        if n_steps // n_interval > 0:
            trajectory.export_mdtraj_h5(prefix, prefix + "_output.h5")    # = self.hdf5_loc when workfolder is the default
        return prefix


class Multi_simulator(Simulator):
    """`num_rep` replicas of one molecule in one resident system plus the single-copy `_ref_system`
    (Simulator.py:1522-1711)."""

    def __init__(self, work_dir: str = None, name: str = "", pdb_id: str = "", mol: Molecule = None, state_dict=None,
                 solvent_model: Sequence[int] = (0,), solvent_dielectric: Sequence[float] = (78.5,), num_rep: int = 1,
                 run_name: str = "", save_name: Optional[str] = None, cache=None, random_number_seed: int = 0,
                 partial_charges=None, rdkit_mol=None, openff_forcefield: str = "openff-2.0.0", constraints=HBonds,
                 device=0, model: str = "gnn"):
        mol = mol if mol is not None else rdkit_mol
        super().__init__(work_dir, name, pdb_id, mol, state_dict, solvent_model, solvent_dielectric, num_rep, run_name,
                         save_name, random_number_seed, constraints, device, model=model)
        self._replicates_exist = False
        self._cache, self._partial_charges, self._openff_forcefield = cache, partial_charges, openff_forcefield
        self.create_ref_system(run_name)

    def create_ref_system(self, run_name):
        """Single-copy system with the first replica's solvent (Simulator.py:1568-1585; create_gnn_sim builds its
        reference model with `solvent_models=[solvent_model[0]]`, helper_functions.py:764-771)."""
        self._ref_system = Simulator(self._work_dir, pdb_id=self._pdb_id, mol=self._mol, state_dict=self._state_dict,
                                     solvent_model=[self._solvent_model[0]], solvent_dielectric=[self._solvent_dielectric[0]],
                                     num_rep=1, run_name=run_name + "ref", random_number_seed=self._random_number_seed,
                                     constraints=self._constraints, device=self._device, model=self._model)

    def setup_replicates(self, only_check_first=False):
        """The reference replicates the OpenMM system and then asserts that every replica feels the forces of the
        single-copy system (Simulator.py:1587-1657: relative difference > 1 % AND absolute > 5 kJ/mol/nm fails).
        Replication itself is implicit here (the engine is built for num_rep copies); the self-check is kept."""
        n = self._mol.GetNumAtoms()
        single = torch.as_tensor(self._mol.conformers_nm[0], dtype=torch.float32)
        self._ref_system.set_positions(single)
        ref_single_forces = self._ref_system.calculate_forces()                       # [3, N]
        self.set_positions(single.unsqueeze(0).expand(self._num_rep, n, 3).contiguous())
        new_forces = self.calculate_forces()                                          # [3, num_rep * N]
        same_solvent = all(s == self._solvent_model[0] and d == self._solvent_dielectric[0]
                           for s, d in zip(self._solvent_model, self._solvent_dielectric))
        for t in range(self._num_rep):
            if not same_solvent and (self._solvent_model[t] != self._solvent_model[0]):
                continue     # a replica in another solvent is not expected to match the reference copy
            diff = np.abs(ref_single_forces - new_forces[:, t * n:(t + 1) * n])
            reldiff = np.abs(diff / ref_single_forces)
            for w in np.argwhere(reldiff > 0.01):
                if np.abs(diff[w[0], w[1]]) > 5:
                    raise AssertionError("Forces are not equal")
            if only_check_first:
                break
        self._replicates_exist = True

    def set_random_positions_for_each_replicate(self, sigma_nm: float = 0.01):
        """The reference embeds num_rep RDKit conformers (Simulator.py:1659-1675); without RDKit the replicas take the
        molecule's own conformers in turn (cycled), plus a seeded Gaussian jitter when there are fewer than num_rep."""
        C, n = self._mol.GetNumConformers(), self._mol.GetNumAtoms()
        rng = np.random.default_rng(self._random_number_seed)
        pos = np.stack([self._mol.conformers_nm[i % C] for i in range(self._num_rep)])
        if C < self._num_rep:
            pos[C:] += rng.normal(0.0, sigma_nm, size=pos[C:].shape)
        self.set_positions(pos.reshape(-1, 3))

    def run_ref_simulation(self, n_interval, n_steps, minimize=False):
        return self._ref_system.run_simulation(n_interval=n_interval, n_steps=n_steps, minimize=minimize)
