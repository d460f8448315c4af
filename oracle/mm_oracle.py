"""CPU oracle of the vacuum molecular-mechanics terms (TEST INFRASTRUCTURE ONLY -- see oracle/gnn_oracle.py).

Restates, with torch autograd for the forces, the energy expressions the reference hands to OpenMM:
  Custom_electrostatic                 charge1 * charge2 * fourpieps / r, fourpieps = 138.935456   Simulation/Simulator.py:2036-2047
  Custom_lennard_jones                 4 sqrt(e1 e2) ((s/r)^12 - (s/r)^6), s = (s1 + s2) / 2        Simulation/Simulator.py:2118-2133
  Custom_exception_force_with_scale    4 eps (sr6^2 - sr6) + chargeprod * fourpieps / r             Simulation/Simulator.py:2086-2103
  exclusions = all NonbondedForce exceptions                                                        Simulation/Simulator.py:2070-2074, 2139-2143
and the three standard OpenMM bonded forces an OpenFF system carries (HarmonicBondForce 0.5 k (r - r0)^2,
HarmonicAngleForce 0.5 k (theta - theta0)^2, PeriodicTorsionForce k (1 + cos(n phi - phase))).

Parity UNPINNED against OpenMM itself: OpenMM is absent here and on the GPU boxes; the dihedral
follows the IUPAC convention OpenMM documents.  The bonded formulas are the published OpenMM definitions."""
import numpy as np
import torch

FOURPIEPS = 138.935456


def dihedral(x0, x1, x2, x3):
    b1, b2, b3 = x1 - x0, x2 - x1, x3 - x2
    n1, n2 = torch.cross(b1, b2, dim=-1), torch.cross(b2, b3, dim=-1)
    lb2 = b2.norm(dim=-1)
    return torch.atan2(lb2 * (b1 * n2).sum(-1), (n1 * n2).sum(-1))


def evaluate(ff, pos, dtype=torch.float64, forces=True):
    """pos [R,N,3]; ff as built by synth.synth_forcefield.  Returns dict(energy_rep [R], forces [R,N,3], parts)."""
    x = torch.as_tensor(np.asarray(pos), dtype=dtype).clone().requires_grad_(forces)
    R, N, _ = x.shape
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
    li = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64)
    parts = {}
    b = ff["bonds"]
    bi = li(b["idx"])
    r = (x[:, bi[:, 0]] - x[:, bi[:, 1]]).norm(dim=-1)
    parts["bond"] = (0.5 * t(b["k"]) * (r - t(b["r0"])) ** 2).sum(-1)
    a = ff["angles"]
    ai = li(a["idx"])
    u, v = x[:, ai[:, 0]] - x[:, ai[:, 1]], x[:, ai[:, 2]] - x[:, ai[:, 1]]
    c = (u * v).sum(-1) / (u.norm(dim=-1) * v.norm(dim=-1))
    theta = torch.acos(c.clamp(-1.0, 1.0))
    parts["angle"] = (0.5 * t(a["k"]) * (theta - t(a["theta0"])) ** 2).sum(-1)
    tt = ff["torsions"]
    ti = li(tt["idx"])
    phi = dihedral(x[:, ti[:, 0]], x[:, ti[:, 1]], x[:, ti[:, 2]], x[:, ti[:, 3]])
    parts["torsion"] = (t(tt["k"]) * (1.0 + torch.cos(t(tt["periodicity"]) * phi - t(tt["phase"])))).sum(-1)
    q, sig, eps = t(ff["charge"]), t(ff["sigma"]), t(ff["epsilon"])
    xc = ff["exceptions"]
    xi = li(xc["idx"])
    mask = torch.ones(N, N, dtype=torch.bool)
    mask[torch.arange(N), torch.arange(N)] = False
    mask[xi[:, 0], xi[:, 1]] = False
    mask[xi[:, 1], xi[:, 0]] = False
    iu = torch.triu_indices(N, N, 1)
    keep = mask[iu[0], iu[1]]
    pi, pj = iu[0][keep], iu[1][keep]
    rij = (x[:, pi] - x[:, pj]).norm(dim=-1)
    parts["coulomb"] = (q[pi] * q[pj] * FOURPIEPS / rij).sum(-1)
    s6 = (0.5 * (sig[pi] + sig[pj]) / rij) ** 6
    parts["lj"] = (4.0 * torch.sqrt(eps[pi] * eps[pj]) * (s6 * s6 - s6)).sum(-1)
    rx = (x[:, xi[:, 0]] - x[:, xi[:, 1]]).norm(dim=-1)
    s6 = (t(xc["sigma"]) / rx) ** 6
    parts["exception"] = (4.0 * t(xc["epsilon"]) * (s6 * s6 - s6) + t(xc["chargeprod"]) * FOURPIEPS / rx).sum(-1)
    e = sum(parts.values())
    out = {"energy_rep": e.detach(), "parts": {k: v.detach() for k, v in parts.items()}}
    if forces:
        (g,) = torch.autograd.grad(e.sum(), x)
        out["forces"] = -g
    return out
