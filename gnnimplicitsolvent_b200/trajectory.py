"""Trajectory output and per-frame energy read-back around the resident batched integrator.

Reference behaviour that is mirrored here:

* `Simulator.run_simulation(n_steps, n_interval, minimize)` (`Simulation/Simulator.py:341-402`): optional
  minimisation, then `simulation.step(n_steps)` with an `HDF5Reporter` (coordinates every `n_interval` steps) and a
  `StateDataReporter` (step, potential energy, temperature, speed) -> `run_simulation` below.
* The replicated system is ONE topology with `num_reps` copies of the molecule, one chain per copy; analysis reads it
  back with `load_parallel_traj(file, nrep)` = `traj.atom_slice(top.select('chainid i'))`
  (`Analysis/analysis.py:2032-2037`) -> same function name and argument order here, same replica-as-chain layout:
  frame f, atom r * N + a = atom a of replica r.
* `get_energies(sim, traj)` (`Simulation/helper_functions.py:901-918`) re-evaluates a trajectory frame by frame on a
  one-replica system -> `get_energies(engine, xyz)` evaluates all frames as replicas of one batched call.

File format.  While a run is in progress the frames go to plain numpy files (memory-mapped, so that device-to-host copies
land in the file directly) that hold exactly the datasets of an MDTraj HDF5 file with the same names, shapes and units:

    <prefix>_output.coordinates.npy   float32 [n_frames, R * N, 3]  nanometers   (memory-mapped while it is written)
    <prefix>_output.meta.npz          time [n_frames] picoseconds, potentialEnergy [n_frames, R] kJ/mol,
                                      temperature [n_frames, R] kelvin, chainid [R * N], element [N], n_replicas, n_atoms
    <prefix>_log.txt                  the StateDataReporter columns

`export_mdtraj_h5` converts them to the MDTraj-convention HDF5 file `<prefix>_output.h5` that the reference's
`HDF5Reporter` writes and `mdtraj.load` reads (h5py where present; here the bytes come from `hdf5_min.py`, a writer for the
"earliest" HDF5 structures -- unpinned against libhdf5, which does not exist in this image).
"""
from __future__ import annotations

import json
import os
import time
from typing import List, Optional, Sequence

import numpy as np
import torch

KB_KJ_MOL_K = 0.00831446261815324   # kJ / (mol K)


class TrajectoryWriter:
    """Fixed-length trajectory of R replicas x N atoms; frames are appended in order."""

    def __init__(self, prefix: str, n_frames: int, n_replicas: int, n_atoms: int,
                 elements: Optional[Sequence[str]] = None):
        self.prefix = prefix
        self.n_frames, self.R, self.N = int(n_frames), int(n_replicas), int(n_atoms)
        self.elements = [str(e) for e in elements] if elements is not None else ["X"] * self.N
        assert len(self.elements) == self.N
        d = os.path.dirname(os.path.abspath(prefix))
        os.makedirs(d, exist_ok=True)
        self.xyz = np.lib.format.open_memmap(self.coordinates_path(prefix), mode="w+", dtype=np.float32,
                                             shape=(self.n_frames, self.R * self.N, 3))
        self.time = np.zeros(self.n_frames, np.float64)
        self.epot = np.full((self.n_frames, self.R), np.nan, np.float64)
        self.temp = np.full((self.n_frames, self.R), np.nan, np.float64)
        self.count = 0
        self._log = open(prefix + "_log.txt", "w")
        self._log.write('#"Step","Potential Energy (kJ/mole)","Temperature (K)","Speed (ns/day)"\n')

    @staticmethod
    def coordinates_path(prefix: str) -> str:
        return prefix + "_output.coordinates.npy"

    @staticmethod
    def meta_path(prefix: str) -> str:
        return prefix + "_output.meta.npz"

    def frame_buffer(self, i: int) -> np.ndarray:
        """The [R, N, 3] slot of frame i inside the memory-mapped file (device-to-host copies land here directly)."""
        return self.xyz[i].reshape(self.R, self.N, 3)

    def append(self, positions, time_ps: float, step: int, potential=None, temperature=None, ns_per_day: float = 0.0):
        i = self.count
        if i >= self.n_frames:
            raise IndexError("trajectory is full")
        if positions is not None:   # None: the caller has filled frame_buffer(i) itself
            p = positions.detach().cpu().numpy() if isinstance(positions, torch.Tensor) else np.asarray(positions)
            self.frame_buffer(i)[...] = p.reshape(self.R, self.N, 3)
        self.time[i] = time_ps
        if potential is not None:
            self.epot[i] = np.asarray(potential, np.float64).reshape(self.R)
        if temperature is not None:
            self.temp[i] = np.asarray(temperature, np.float64).reshape(self.R)
        t_mean = float(np.nanmean(self.temp[i])) if temperature is not None else float("nan")
        self._log.write(f"{step},{np.nansum(self.epot[i]):.6f},{t_mean:.4f},{ns_per_day:.3f}\n")
        self.count += 1

    def close(self):
        if self._log is None:
            return
        self.xyz.flush()
        del self.xyz
        np.savez(self.meta_path(self.prefix), time=self.time[: self.count], potentialEnergy=self.epot[: self.count],
                 temperature=self.temp[: self.count], chainid=np.repeat(np.arange(self.R), self.N),
                 element=np.asarray(self.elements), n_replicas=self.R, n_atoms=self.N, n_frames=self.count)
        self._log.close()
        self._log = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def load_trajectory(prefix: str):
    """-> (coordinates [frames, R*N, 3] memmap, meta dict)."""
    meta = dict(np.load(TrajectoryWriter.meta_path(prefix), allow_pickle=False))
    xyz = np.load(TrajectoryWriter.coordinates_path(prefix), mmap_mode="r")
    return xyz[: int(meta["n_frames"])], meta


def load_parallel_traj(file: str, nrep: int) -> List[np.ndarray]:
    """One [frames, N, 3] array per replica (chain), as `Analysis/analysis.py:2032-2037` returns one trajectory per
    `chainid`.  `file` is the prefix given to TrajectoryWriter, or the `.h5` file export_mdtraj_h5 wrote from it (selected
    by chain id, as `traj.top.select('chainid %i')` does)."""
    if file.endswith(".h5"):
        xyz, _, chainid, _ = load_mdtraj_h5(file)
        if nrep > int(chainid.max()) + 1:
            raise ValueError(f"trajectory holds {int(chainid.max()) + 1} replicas, {nrep} requested")
        return [xyz[:, chainid == r, :] for r in range(nrep)]
    xyz, meta = load_trajectory(file)
    R, N = int(meta["n_replicas"]), int(meta["n_atoms"])
    if nrep > R:
        raise ValueError(f"trajectory holds {R} replicas, {nrep} requested")
    return [xyz[:, r * N:(r + 1) * N, :] for r in range(nrep)]


def kinetic_temperature(velocities: torch.Tensor, masses, n_constraints: int = 0) -> torch.Tensor:
    """Instantaneous temperature per replica [R] (kelvin): 2 KE / (dof kB), dof = 3N - constraints (OpenMM's
    StateDataReporter additionally subtracts 3 when a CMMotionRemover is present; the reference system has none)."""
    m = torch.as_tensor(np.asarray(masses, np.float32), device=velocities.device).reshape(1, -1, 1)
    ke = 0.5 * (m * velocities.double() ** 2).sum(dim=(1, 2))
    dof = 3 * velocities.shape[1] - int(n_constraints)
    return 2.0 * ke / (dof * KB_KJ_MOL_K)


def run_simulation(engine, positions: torch.Tensor, velocities: torch.Tensor, masses, n_steps: int = 10000,
                   n_interval: int = 1000, minimize: bool = True, prefix: Optional[str] = None, *,
                   temperature: float = 300.0, friction: float = 1.0, dt: float = 0.002, seed: int = 0,
                   elements: Optional[Sequence[str]] = None, n_constraints: int = 0, grad_tol: float = 10.0,
                   max_minimize_rounds: int = 500, first_step: int = 0, eval_flags: Optional[dict] = None):
    """`Simulator.run_simulation` for the resident batched system: optional minimisation, then n_steps of
    LangevinMiddle in blocks of n_interval steps; after every block the frame is copied to the host on a side stream
    (into the memory-mapped file, overlapping the next block) and one log line is written.

    positions / velocities: CUDA [R, N, 3], advanced in place.  Returns the TrajectoryWriter prefix (or None)."""
    assert positions.is_cuda and velocities.is_cuda
    R, N = positions.shape[0], positions.shape[1]
    eval_flags = dict(eval_flags or {})     # Hamiltonian switches of the engine (gb_only= / vacuum_only=)
    if minimize:
        pmin, _, _, _ = engine.minimize(positions, grad_tol=grad_tol, max_iter=max_minimize_rounds, **eval_flags)
        positions.copy_(pmin)
    n_frames = n_steps // n_interval
    writer = TrajectoryWriter(prefix, n_frames, R, N, elements) if prefix is not None else None
    copy_stream = torch.cuda.Stream(device=positions.device)
    # two device snapshots + two pinned staging buffers: the copy of frame i runs under the dynamics of block i + 1
    snaps = [torch.empty_like(positions) for _ in range(2)] if writer is not None else []
    staging = [torch.empty((R, N, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)] if writer is not None else []
    pending = None   # (frame index, staging slot, event, meta)
    done = int(first_step)   # the random stream is keyed by the absolute step number: continuing runs do not repeat it
    t_wall = time.perf_counter()

    def flush(p):
        i, slot, ev, meta = p
        ev.synchronize()
        writer.frame_buffer(i)[...] = staging[slot].numpy()
        writer.append(None, **meta)

    for i in range(n_frames):
        engine.langevin(positions, velocities, masses, n_interval, temperature=temperature, friction=friction, dt=dt,
                        seed=seed, first_step=done, **eval_flags)
        done += n_interval
        if writer is None:
            continue
        slot = i & 1   # its previous user (frame i - 2) was flushed in iteration i - 1
        e, _ = engine.energy_force(positions, forces=False, **eval_flags)
        temp = kinetic_temperature(velocities, masses, n_constraints)
        snaps[slot].copy_(positions)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(positions.device))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            staging[slot].copy_(snaps[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        if pending is not None:
            flush(pending)
        elapsed = max(time.perf_counter() - t_wall, 1e-9)
        meta = dict(time_ps=done * dt, step=done, potential=e.double().cpu().numpy(), temperature=temp.cpu().numpy(),
                    ns_per_day=done * dt * 1e-3 / elapsed * 86400.0)
        pending = (i, slot, ev, meta)
    if pending is not None:
        flush(pending)
    if writer is not None:
        writer.close()
    return prefix


def get_energies(engine, xyz, solvent_id: int = 0, dielectric: float = 78.5, batch: int = 4096, **flags) -> np.ndarray:
    """Potential energy of every frame of `xyz` [frames, N, 3] (kJ/mol): the frames are evaluated as replicas of
    batched calls instead of the reference's one-frame-at-a-time loop (`helper_functions.py:901-918`).  The engine's
    replica set-up is replaced; frames that fail to evaluate come back as NaN, as in the reference."""
    x = np.asarray(xyz, np.float32)
    F = x.shape[0]
    out = np.full(F, np.nan, np.float64)
    for lo in range(0, F, batch):
        hi = min(F, lo + batch)
        engine.set_replicas(hi - lo, [solvent_id] * (hi - lo), [dielectric] * (hi - lo))
        e, _ = engine.energy_force(torch.from_numpy(np.ascontiguousarray(x[lo:hi])).to(engine.device), forces=False,
                                   **flags)
        out[lo:hi] = e.double().cpu().numpy()
    return out


def _mdtraj_topology(R: int, N: int, elements, residue_name: str = "MOL") -> dict:
    """MDTraj's JSON topology: one chain (and one residue) per replica, as the reference's replicated OpenMM system."""
    chains, idx = [], 0
    for r in range(R):
        atoms = []
        for a in range(N):
            atoms.append({"index": idx, "name": f"{elements[a]}{a + 1}", "element": elements[a]})
            idx += 1
        chains.append({"index": r, "residues": [{"index": r, "name": residue_name, "resSeq": 1, "atoms": atoms}]})
    return {"chains": chains, "bonds": []}


def export_mdtraj_h5(prefix: str, out_file: str, residue_name: str = "MOL"):
    """Writes the MDTraj-convention HDF5 file the reference's HDF5Reporter produces (`Simulation/Simulator.py:364-401`:
    `coordinates` nm, `time` ps, `potentialEnergy` kJ/mol, `topology` JSON with one chain per replica, the Pande-convention
    root attributes) from a TrajectoryWriter output.  Uses h5py where it exists; in this image (no h5py / PyTables /
    libhdf5) the bytes are laid out by `hdf5_min.write` (contiguous datasets, "earliest" HDF5 structures)."""
    xyz, meta = load_trajectory(prefix)
    R, N = int(meta["n_replicas"]), int(meta["n_atoms"])
    topology = json.dumps(_mdtraj_topology(R, N, [str(e) for e in meta["element"]], residue_name)).encode("ascii")
    root = {"conventions": "Pande", "conventionVersion": "1.1", "program": "gnnimplicitsolvent_b200", "programVersion": "0.1",
            "application": "gnnis", "title": os.path.basename(prefix)}
    data = {"coordinates": xyz, "time": meta["time"].astype(np.float32),   # xyz: float32 memmap, written without a copy
            "potentialEnergy": np.nansum(meta["potentialEnergy"], axis=1).astype(np.float32),
            "topology": np.array([topology], dtype=f"S{len(topology)}")}
    units = {"coordinates": {"units": "nanometers"}, "time": {"units": "picoseconds"},
             "potentialEnergy": {"units": "kilojoules_per_mole"}}
    try:
        import h5py  # noqa: PLC0415  (optional dependency)
    except ImportError:
        from . import hdf5_min
        hdf5_min.write(out_file, data, root, units)
        return out_file
    with h5py.File(out_file, "w") as f:
        for k, v in root.items():
            f.attrs[k] = v
        for k, v in data.items():
            d = f.create_dataset(k, data=v)
            for ak, av in units.get(k, {}).items():
                d.attrs[ak] = av
    return out_file


def load_mdtraj_h5(file: str):
    """-> (coordinates [frames, atoms, 3] nm, time [frames] ps, chainid [atoms], elements [atoms]) of an MDTraj-convention
    HDF5 file written by export_mdtraj_h5 (h5py where present, hdf5_min.read otherwise)."""
    try:
        import h5py  # noqa: PLC0415
        with h5py.File(file, "r") as f:
            data = {k: f[k][...] for k in ("coordinates", "time", "topology")}
    except ImportError:
        from . import hdf5_min
        data, _, _ = hdf5_min.read(file)
    top = json.loads(bytes(data["topology"].reshape(-1)[0]).decode("ascii"))
    chainid, elements = [], []
    for ch in top["chains"]:
        for res in ch["residues"]:
            for at in res["atoms"]:
                chainid.append(ch["index"])
                elements.append(at["element"])
    return data["coordinates"], data["time"], np.asarray(chainid), elements
