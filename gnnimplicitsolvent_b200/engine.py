"""Engine: thin Python owner of one libgnnis handle (one per process / GPU).

PyTorch is used only as plumbing: device tensors (data_ptr), the current CUDA stream and pinned host
buffers.  Every numerical result comes from the hand-written sm_100a kernels behind include/gnnis.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, gbneck

_LAYER_KEYS = ("message1.weight", "message1.bias", "message2.weight", "message2.bias",
               "lin1.weight", "lin1.bias", "lin2.weight", "lin2.bias")


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_float_p)


class GnnisError(RuntimeError):
    pass


class Engine:
    """Batched GB-Neck2 + multisolvent-GNN energy/force engine for R replicas of one N-atom molecule.

    parameters: [N,7] table [q, rho, s, alpha, beta, gamma, radindex] (MachineLearning/GNN_Trainer.py:829-869)
    state_dict: the 39-key state dict of the shipped architecture (or None for GB-only use)
    """

    def __init__(self, parameters, state_dict=None, device: int | torch.device | str = 0,
                 unique_radii=gbneck.DEFAULT_UNIQUE_RADII, fraction: float = 0.1, scaling_factor: float = 2.0,
                 cutoff: float = 0.6):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise GnnisError("no CUDA device: the gnnis hot path has no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise GnnisError(f"Engine needs a CUDA device, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._h = C.c_void_p()
        if self.lib.gnnis_create(C.byref(self._h), self.device.index) != 0:
            raise GnnisError(self.lib.gnnis_last_error(None).decode())
        self.n_rep = 0
        self.n_solvents = 0
        self._check(self.lib.gnnis_set_model_constants(self._h, float(fraction), float(scaling_factor), float(cutoff)))
        if state_dict is not None:
            self.load_state_dict(state_dict)
        params = _f32(parameters)
        if params.ndim != 2 or params.shape[1] != 7:
            raise ValueError("parameters must be [N,7]")
        self.n_atoms = int(params.shape[0])
        d0 = m0 = None
        if state_dict is not None and "aggregate_information._d0" in state_dict:
            d0, m0 = _f32(state_dict["aggregate_information._d0"]), _f32(state_dict["aggregate_information._m0"])
            n_unique = int(round(np.sqrt(d0.size)))
        if d0 is None:
            d0, m0, n_unique = gbneck.neck_tables(unique_radii)
        self._check(self.lib.gnnis_set_system(self._h, self.n_atoms, _fp(params), _fp(d0), _fp(m0), n_unique))
        self._host = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise GnnisError(f"libgnnis error {rc}: {self.lib.gnnis_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.gnnis_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _pos(self, positions: torch.Tensor) -> torch.Tensor:
        if positions.device != self.device:
            positions = positions.to(self.device)
        p = positions.detach().to(torch.float32).reshape(self.n_rep, self.n_atoms, 3).contiguous()
        return p

    # ------------------------------------------------------------------ configuration
    def load_state_dict(self, sd):
        w = _lib.GnnisWeights()
        keep = []
        emb = _f32(sd["solvent_embedding.weight"])
        gam = _f32(sd["gamma_embedding.weight"]).reshape(-1)
        keep += [emb, gam]
        w.hidden, w.hidden_token, w.num_solvents, w.n_rbf = emb.shape[1], _f32(sd["interaction1.lin1.weight"]).shape[0], emb.shape[0], 20
        w.solvent_embedding, w.gamma_embedding = _fp(emb), _fp(gam)
        for l in range(3):
            arrs = [_f32(sd[f"interaction{l + 1}.{k}"]) for k in _LAYER_KEYS]
            keep += arrs
            (w.message1_w[l], w.message1_b[l], w.message2_w[l], w.message2_b[l],
             w.lin1_w[l], w.lin1_b[l], w.lin2_w[l], w.lin2_b[l]) = [_fp(a) for a in arrs]
        self._check(self.lib.gnnis_load_weights(self._h, C.byref(w)))
        self.n_solvents = int(emb.shape[0])

    def set_replicas(self, n_rep: int, solvent_ids: Sequence[int], dielectrics: Sequence[float]):
        """set_num_reps(num_reps, solvent_models, solvent_dielectric) (GNN_Models.py:741-769)."""
        sid = np.ascontiguousarray(np.asarray(solvent_ids, dtype=np.int32).reshape(-1))
        die = np.ascontiguousarray(np.asarray(dielectrics, dtype=np.float32).reshape(-1))
        if sid.size != n_rep or die.size != n_rep:
            raise ValueError("need one solvent id and one dielectric per replica")
        self._check(self.lib.gnnis_set_replicas(self._h, int(n_rep), sid.ctypes.data_as(_lib.c_int32_p), _fp(die)))
        self.n_rep = int(n_rep)
        self._host = None

    def set_forcefield(self, ff: Optional[dict]):
        """Vacuum MM terms: dict with `bonds` (idx [B,2], r0, k), `angles` (idx [A,3], theta0, k),
        `torsions` (idx [T,4], periodicity, phase, k), `charge`, `sigma`, `epsilon` [N] and `exceptions`
        (idx [X,2], chargeprod, sigma, epsilon) -- the parameters of an OpenMM System's HarmonicBond / HarmonicAngle /
        PeriodicTorsion / NonbondedForce.  None removes the force field."""
        if ff is None:
            self._check(self.lib.gnnis_set_forcefield(self._h, None))
            return
        s = _lib.GnnisForcefield()
        keep = []

        def ints(a, w):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1, w))
            keep.append(a)
            return a, a.ctypes.data_as(_lib.c_int32_p)

        def flt(a):
            a = _f32(a).reshape(-1)
            keep.append(a)
            return _fp(a)

        b = ff.get("bonds", {"idx": np.zeros((0, 2)), "r0": [], "k": []})
        a = ff.get("angles", {"idx": np.zeros((0, 3)), "theta0": [], "k": []})
        t = ff.get("torsions", {"idx": np.zeros((0, 4)), "periodicity": [], "phase": [], "k": []})
        x = ff.get("exceptions", {"idx": np.zeros((0, 2)), "chargeprod": [], "sigma": [], "epsilon": []})
        bi, s.bond_idx = ints(b["idx"], 2)
        s.n_bonds, s.bond_r0, s.bond_k = bi.shape[0], flt(b["r0"]), flt(b["k"])
        ai, s.angle_idx = ints(a["idx"], 3)
        s.n_angles, s.angle_theta0, s.angle_k = ai.shape[0], flt(a["theta0"]), flt(a["k"])
        ti, s.torsion_idx = ints(t["idx"], 4)
        per = np.ascontiguousarray(np.asarray(t["periodicity"], dtype=np.int32).reshape(-1))
        keep.append(per)
        s.n_torsions, s.torsion_periodicity = ti.shape[0], per.ctypes.data_as(_lib.c_int32_p)
        s.torsion_phase, s.torsion_k = flt(t["phase"]), flt(t["k"])
        for name in ("charge", "sigma", "epsilon"):
            arr = _f32(ff[name]).reshape(-1)
            if arr.size != self.n_atoms:
                raise ValueError(f"forcefield['{name}'] needs one value per atom")
            keep.append(arr)
            setattr(s, name, _fp(arr))
        xi, s.exception_idx = ints(x["idx"], 2)
        s.n_exceptions = xi.shape[0]
        s.exception_chargeprod, s.exception_sigma, s.exception_epsilon = flt(x["chargeprod"]), flt(x["sigma"]), flt(x["epsilon"])
        self._check(self.lib.gnnis_set_forcefield(self._h, C.byref(s)))

    def set_constraints(self, idx, dist, tolerance: float = 1e-5):
        """Distance constraints for `langevin` (e.g. all X-H bonds at their force-field length = OpenMM's HBonds)."""
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int32).reshape(-1, 2))
        dist = _f32(dist).reshape(-1)
        self._check(self.lib.gnnis_set_constraints(self._h, idx.shape[0], idx.ctypes.data_as(_lib.c_int32_p), _fp(dist),
                                                   float(tolerance)))

    def set_bonds(self, idx, r0, k):
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int32).reshape(-1, 2))
        r0, k = _f32(r0).reshape(-1), _f32(k).reshape(-1)
        self._check(self.lib.gnnis_set_bonds(self._h, idx.shape[0], idx.ctypes.data_as(_lib.c_int32_p), _fp(r0), _fp(k)))

    # ------------------------------------------------------------------ hot path
    def energy_force(self, positions: torch.Tensor, *, delta=False, forces=True, gb_only=False, cuda_cores=False,
                     data_predicate=False, vacuum_only=False, tf32x1=False, bf16=False, fp16=False, cell_list=False, extra_flags: int = 0,
                     out_energy: Optional[torch.Tensor] = None, out_force: Optional[torch.Tensor] = None):
        """positions [R,N,3] (or [R*N,3]) fp32 on the engine's device -> (E[R], F[R,N,3] or None)."""
        p = self._pos(positions)
        e = out_energy if out_energy is not None else torch.empty(self.n_rep, dtype=torch.float32, device=self.device)
        f = None
        if forces:
            f = out_force if out_force is not None else torch.empty_like(p)
        flags = ((_lib.FLAG_DELTA if delta else 0) | (0 if forces else _lib.FLAG_NO_FORCES)
                 | (_lib.FLAG_GB_ONLY if gb_only else 0) | (_lib.FLAG_MLP_CUDA_CORES if cuda_cores else 0)
                 | (_lib.FLAG_DATA_PREDICATE if data_predicate else 0) | (_lib.FLAG_VACUUM_ONLY if vacuum_only else 0)
                 | (_lib.FLAG_MLP_TF32X1 if tf32x1 else 0) | (_lib.FLAG_MLP_BF16 if bf16 else 0) | (_lib.FLAG_MLP_FP16 if fp16 else 0)
                 | (_lib.FLAG_CELL_LIST if cell_list else 0) | int(extra_flags))
        self._check(self.lib.gnnis_energy_force(self._h, p.data_ptr(), e.data_ptr(), f.data_ptr() if forces else None,
                                                flags, self._stream()))
        return e, f

    def energy_force_host(self, pos_host: torch.Tensor, e_host: torch.Tensor, f_host: Optional[torch.Tensor], *,
                          delta=False, gb_only=False):
        """End-to-end call with HOST tensors (pinned recommended): H2D + evaluation + D2H inside libgnnis."""
        flags = (_lib.FLAG_DELTA if delta else 0) | (_lib.FLAG_GB_ONLY if gb_only else 0) | (0 if f_host is not None else _lib.FLAG_NO_FORCES)
        assert pos_host.dtype == torch.float32 and pos_host.is_contiguous() and not pos_host.is_cuda
        self._check(self.lib.gnnis_energy_force_host(self._h, pos_host.data_ptr(), e_host.data_ptr(),
                                                     f_host.data_ptr() if f_host is not None else None, flags))
        return e_host, f_host

    def energy_force_host_submit(self, pos_host: torch.Tensor, e_host: torch.Tensor, f_host: Optional[torch.Tensor], *,
                                 delta=False, gb_only=False):
        """Pipelined form of energy_force_host: enqueues H2D + evaluation + D2H and returns at once (at most two
        submissions in flight).  The tensors must stay alive and unread until the matching energy_force_host_wait()."""
        flags = (_lib.FLAG_DELTA if delta else 0) | (_lib.FLAG_GB_ONLY if gb_only else 0) | (0 if f_host is not None else _lib.FLAG_NO_FORCES)
        assert pos_host.dtype == torch.float32 and pos_host.is_contiguous() and not pos_host.is_cuda
        self._check(self.lib.gnnis_energy_force_host_submit(self._h, pos_host.data_ptr(), e_host.data_ptr(),
                                                            f_host.data_ptr() if f_host is not None else None, flags))

    def energy_force_host_wait(self):
        """Blocks until the OLDEST submission in flight has delivered its energies / forces."""
        self._check(self.lib.gnnis_energy_force_host_wait(self._h))

    def atom_energies(self):
        """Per-atom polar and SA energies of the last evaluation (..._split, GNN_Models.py:1068-1083)."""
        n = self.n_rep * self.n_atoms
        e = torch.empty(n, dtype=torch.float32, device=self.device)
        sa = torch.empty(n, dtype=torch.float32, device=self.device)
        self._check(self.lib.gnnis_get_atom_energies(self._h, e.data_ptr(), sa.data_ptr(), self._stream()))
        return e, sa

    def born_radii(self):
        n = self.n_rep * self.n_atoms
        b = torch.empty(n, dtype=torch.float32, device=self.device)
        bs = torch.empty(n, dtype=torch.float32, device=self.device)
        self._check(self.lib.gnnis_get_born_radii(self._h, b.data_ptr(), bs.data_ptr(), self._stream()))
        return b, bs

    def last_edge_count(self) -> int:
        v = C.c_int64(0)
        self._check(self.lib.gnnis_last_edge_count(self._h, C.byref(v), self._stream()))
        return int(v.value)

    def build_neighbors(self, positions: torch.Tensor, predicate: str = "run", backend: Optional[str] = None):
        """Sorted CSR radius graph (target-major, ascending source): (row_ptr[n+1], col[E], r[E]).  backend: "cell" (cell
        list), "allpairs", or None = the handle's default; both give the same bytes."""
        p = self._pos(positions)
        n = self.n_rep * self.n_atoms
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=self.device)
        pred = (0 if predicate == "run" else 1) | {None: 0, "cell": _lib.PREDICATE_CELL_LIST, "allpairs": _lib.PREDICATE_ALL_PAIRS}[backend]
        ne = C.c_int64(0)
        cap = max(1, n * min(self.n_atoms - 1, 64))
        while True:
            col = torch.empty(cap, dtype=torch.int32, device=self.device)
            r = torch.empty(cap, dtype=torch.float32, device=self.device)
            rc = self.lib.gnnis_build_neighbors(self._h, p.data_ptr(), pred, row_ptr.data_ptr(), col.data_ptr(),
                                                r.data_ptr(), cap, C.byref(ne), self._stream())
            if rc == -2:
                cap = int(ne.value)
                continue
            self._check(rc)
            break
        e = int(ne.value)
        return row_ptr, col[:e], r[:e]

    def minimize(self, positions: torch.Tensor, grad_tol: float = 1.0, max_iter: int = 500, history: int = 8,
                 delta=False, gb_only=False, vacuum_only=False):
        """Batched per-replica L-BFGS (history = curvature pairs per replica; history=-1: steepest descent with the same line
        search); returns (positions[R,N,3], E[R], status[R] int32, n_rounds)."""
        p = self._pos(positions).clone()
        e = torch.empty(self.n_rep, dtype=torch.float32, device=self.device)
        status = torch.empty(self.n_rep, dtype=torch.int32, device=self.device)
        nit = C.c_int32(0)
        flags = ((_lib.FLAG_DELTA if delta else 0) | (_lib.FLAG_GB_ONLY if gb_only else 0)
                 | (_lib.FLAG_VACUUM_ONLY if vacuum_only else 0))
        self._check(self.lib.gnnis_minimize(self._h, p.data_ptr(), e.data_ptr(), status.data_ptr(), C.byref(nit),
                                            float(grad_tol), int(max_iter), int(history), flags, self._stream()))
        return p, e, status, int(nit.value)

    def minimize_stats(self):
        """(evaluations[R] int32, gradient RMS[R] float32) of the last `minimize` call, per replica."""
        n = torch.empty(self.n_rep, dtype=torch.int32, device=self.device)
        g = torch.empty(self.n_rep, dtype=torch.float32, device=self.device)
        self._check(self.lib.gnnis_minimize_stats(self._h, n.data_ptr(), g.data_ptr(), self._stream()))
        return n, g

    def langevin(self, positions: torch.Tensor, velocities: torch.Tensor, masses, n_steps: int, *, temperature=300.0,
                 friction=1.0, dt=0.002, seed=0, first_step=0, delta=False, gb_only=False, vacuum_only=False):
        """LangevinMiddle steps in place on (positions, velocities) [R,N,3]."""
        assert positions.is_cuda and velocities.is_cuda and positions.is_contiguous() and velocities.is_contiguous()
        m = _f32(masses).reshape(-1)
        flags = ((_lib.FLAG_DELTA if delta else 0) | (_lib.FLAG_GB_ONLY if gb_only else 0)
                 | (_lib.FLAG_VACUUM_ONLY if vacuum_only else 0))
        self._check(self.lib.gnnis_langevin(self._h, positions.data_ptr(), velocities.data_ptr(), _fp(m),
                                            float(temperature), float(friction), float(dt), int(n_steps), int(seed),
                                            int(first_step), flags, self._stream()))

    def umma_selftest(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """D = A @ B^T through the tcgen05 building blocks (A [128,K], B [N,K], K,N in {32,64})."""
        a = a.to(self.device, torch.float32).contiguous()
        b = b.to(self.device, torch.float32).contiguous()
        d = torch.empty(128, b.shape[0], dtype=torch.float32, device=self.device)
        self._check(self.lib.gnnis_umma_selftest(self._h, a.data_ptr(), b.data_ptr(), d.data_ptr(), a.shape[1], b.shape[0],
                                                 self._stream()))
        return d

    def describe(self) -> str:
        """Which code path the current set-up runs on (tcgen05 / CUDA cores, stash / recompute, tables, split)."""
        return self.lib.gnnis_describe(self._h).decode()

    def launch_count(self) -> int:
        return int(self.lib.gnnis_launch_count(self._h))

    def profile_eval(self, positions: torch.Tensor, *, delta=False, extra_flags: int = 0):
        p = self._pos(positions)
        e = torch.empty(self.n_rep, dtype=torch.float32, device=self.device)
        f = torch.empty_like(p)
        ns = self.lib.gnnis_num_stages()
        ms = (C.c_float * ns)()
        self._check(self.lib.gnnis_profile_eval(self._h, p.data_ptr(), e.data_ptr(), f.data_ptr(),
                                                (_lib.FLAG_DELTA if delta else 0) | int(extra_flags), self._stream(), ms, ns))
        return {self.lib.gnnis_stage_name(i).decode(): float(ms[i]) for i in range(ns)}
