"""Batched Hessians, normal modes and quasi-RRHO entropies.

Host-side mirror of the reference's entropy path (Simulation/helper_functions.py):
    build_mass_weighted_hessian  :1551-1622   central differences of the forces, one coordinate at a time
    normal_modes                 :1625-1683   eigh of the mass-weighted Hessian -> wavenumbers, 6 smallest dropped
    calculate_entropy_from_frequencies :1766-1810   Grimme's quasi-RRHO interpolation (T = 300 K)
    calculate_rotational_entropy :1813-1885   rigid rotor from the inertia tensor
    calculate_entropy            :1688-1745   minimise without constraints, then S per conformer; returns E - S
The reference spends 6N sequential single-replica force evaluations per conformer; here all 6N displaced copies of
all C conformers are ONE batched evaluation of the engine (R = 6 N C replicas resident on the GPU).  The
diagonalisation is torch.linalg.eigh in fp64 (library call, not on the hot path)."""
from __future__ import annotations

import numpy as np
import torch

from .engine import Engine

KB = 1.380649e-23        # J/K
H_PLANCK = 6.62607015e-34
N_A = 6.02214076e23
C_CM = 2.99792458e10     # cm/s
DALTON = 1.66054e-27     # kg, the rounded value the reference uses (helper_functions.py:1834)


def mass_weighted_hessians(engine: Engine, positions, masses, solvent_id=0, dielectric=78.5, delta=1e-4,
                           chunk=None, **eval_kw) -> torch.Tensor:
    """positions [C,N,3] (nm) -> mass-weighted Hessians [C,3N,3N] fp64, kJ/(mol nm^2 dalton), symmetrised as the
    reference does (average with the transpose).  `eval_kw` goes to Engine.energy_force (delta=, vacuum_only=, ...)."""
    pos = torch.as_tensor(np.asarray(positions), dtype=torch.float32)
    C, N, _ = pos.shape
    m = torch.as_tensor(np.asarray(masses), dtype=torch.float64).reshape(N)
    D = 3 * N
    per = 2 * D                      # displaced copies per conformer
    chunk = chunk or max(1, min(C, 65536 // per))
    out = torch.empty(C, D, D, dtype=torch.float64, device=engine.device)
    eye = torch.eye(D, dtype=torch.float32).reshape(D, N, 3) * float(delta)
    disp = torch.cat([eye, -eye], 0).to(engine.device)                # [2D, N, 3]
    inv_sqrt_m = (1.0 / torch.sqrt(m)).repeat_interleave(3).to(engine.device)
    for c0 in range(0, C, chunk):
        c1 = min(C, c0 + chunk)
        nc = c1 - c0
        R = nc * per
        if engine.n_rep != R:
            engine.set_replicas(R, [solvent_id] * R, [dielectric] * R)
        batch = (pos[c0:c1].to(engine.device)[:, None] + disp[None]).reshape(R, N, 3).contiguous()
        _, f = engine.energy_force(batch, **eval_kw)
        g = -f.reshape(nc, 2, D, D).double()                          # gradients, [conf, +/-, displaced coord, coord]
        hess = (g[:, 0] - g[:, 1]) / (2.0 * float(delta))
        hess = hess * inv_sqrt_m[None, :, None] * inv_sqrt_m[None, None, :]
        out[c0:c1] = 0.5 * (hess + hess.transpose(1, 2))
    return out


def normal_mode_wavenumbers(hessians: torch.Tensor, n_atoms: int) -> np.ndarray:
    """[C,3N,3N] -> complex wavenumbers [C, 3N-6] (cm^-1; imaginary = negative curvature), sorted by eigenvalue with
    the 6 (5 for diatomics) smallest |nu| removed -- helper_functions.py:1649-1682."""
    eig = torch.linalg.eigvalsh(hessians).cpu().numpy()
    coef = 0.5 / np.pi * 33.3564095
    neg = (eig >= 0).astype(int) * 2 - 1
    freqs = np.sqrt(eig + 0j) * coef * neg
    n_remove = 5 if n_atoms == 2 else 6
    out = []
    for fr in freqs:
        keep = np.sort(np.argpartition(np.abs(fr), n_remove)[n_remove:])
        out.append(fr[keep])
    return np.asarray(out)


def vibrational_entropy(wavenumbers, temperature=300.0) -> float:
    """T*S_vib in kJ/mol from wavenumbers (cm^-1), Grimme's quasi-RRHO -- helper_functions.py:1766-1810."""
    nu = np.asarray(wavenumbers, dtype=np.clongdouble) * C_CM
    T = np.longdouble(temperature)
    kB, h, NA = np.longdouble(KB), np.longdouble(H_PLANCK), np.longdouble(N_A)
    x = h * nu / (kB * T)
    sv = -np.log(1 - np.exp(-x)) + x / (np.exp(x) - 1)
    sv = kB * sv * T * NA / 1000
    bav = 10e-44
    mu = h / (8 * np.pi ** 2 * nu)
    mup = (mu * bav) / (mu + bav)
    sr = 0.5 + np.log(np.sqrt((8 * np.pi ** 3 * mup * kB * T) / h ** 2))
    sr = sr * kB * T * NA / 1000
    w0 = np.longdouble(100 * C_CM)
    w = 1 / (1 + (w0 / nu) ** 4)
    return float(np.sum(w * sv + (1 - w) * sr).real)


def rotational_entropy(positions_nm, masses, temperature=300.0) -> float:
    """T*S_rot in kJ/mol of a rigid rotor -- helper_functions.py:1813-1885 (positions in nm here, metres there)."""
    pos = np.asarray(positions_nm, dtype=np.float64) * 1e-9
    m = np.asarray(masses, dtype=np.longdouble) * np.longdouble(DALTON)
    com = (m[:, None] * pos).sum(0) / m.sum()
    r = pos - np.asarray(com, dtype=np.float64)
    mm = np.asarray(m, dtype=np.float64)
    inertia = np.zeros((3, 3))
    inertia[0, 0] = (mm * (r[:, 1] ** 2 + r[:, 2] ** 2)).sum()
    inertia[1, 1] = (mm * (r[:, 0] ** 2 + r[:, 2] ** 2)).sum()
    inertia[2, 2] = (mm * (r[:, 0] ** 2 + r[:, 1] ** 2)).sum()
    inertia[0, 1] = inertia[1, 0] = -(mm * r[:, 0] * r[:, 1]).sum()
    inertia[1, 2] = inertia[2, 1] = -(mm * r[:, 1] * r[:, 2]).sum()
    inertia[0, 2] = inertia[2, 0] = -(mm * r[:, 0] * r[:, 2]).sum()
    eig = np.linalg.eigvals(inertia)
    kB, h, NA, T = np.longdouble(KB), np.longdouble(H_PLANCK), np.longdouble(N_A), np.longdouble(temperature)
    theta = h ** 2 / (8 * np.pi ** 2 * eig * kB)
    zrot = np.sqrt(np.pi) * np.sqrt(T ** 3 / (theta[0] * theta[1] * theta[2]))
    return float(kB * T * (1.5 + np.log(zrot)) * NA / 1000)


def conformer_entropies(engine: Engine, positions, masses, solvent_id=0, dielectric=78.5, delta=1e-4, **eval_kw):
    """T*S (vibrational quasi-RRHO + rotational) per conformer, kJ/mol, and the wavenumbers.  Conformers whose
    lowest mode is imaginary are reported as they are (the reference re-minimises them, :1757-1759)."""
    pos = np.asarray(positions, dtype=np.float32)
    hess = mass_weighted_hessians(engine, pos, masses, solvent_id, dielectric, delta, **eval_kw)
    nus = normal_mode_wavenumbers(hess, pos.shape[1])
    s = np.array([vibrational_entropy(nu) + rotational_entropy(p, masses) for nu, p in zip(nus, pos)])
    return s, nus


def calculate_entropy(engine: Engine, positions, masses, solvent_id=0, dielectric=78.5, grad_tol=1.0, max_iter=2000,
                      return_free_energy=True, delta=1e-4):
    """calculate_entropy of the reference (helper_functions.py:1688-1745) for conformers given as coordinates:
    minimise every conformer on the engine's Hamiltonian (no constraints), then T*S per conformer from the batched
    Hessians.  Returns (entropies, energies - entropies) like the reference, or entropies alone."""
    pos = torch.as_tensor(np.asarray(positions), dtype=torch.float32)
    C = pos.shape[0]
    engine.set_replicas(C, [solvent_id] * C, [dielectric] * C)
    pmin, e, status, _ = engine.minimize(pos.to(engine.device), grad_tol=grad_tol, max_iter=max_iter)
    s, _ = conformer_entropies(engine, pmin.cpu().numpy(), masses, solvent_id, dielectric, delta)
    engine.set_replicas(C, [solvent_id] * C, [dielectric] * C)
    if return_free_energy:
        return s, e.cpu().numpy().astype(np.float64) - s
    return s
