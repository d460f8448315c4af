"""Size-independent physical properties of the CUDA path at the benchmark's full size (BASELINE.json C2: N = 50 atoms,
R = 4096 conformers).  The oracle cannot be run at this size in seconds; these checks need no reference values:
translation / rotation invariance of the energy, covariance of the forces, vanishing net force and torque, forces
equal to the negative directional derivative of the energy, and independence of a replica from its batch.

The model's energy is discontinuous where a pair crosses the 0.6 nm graph cutoff, so a few of the
4096 conformers may change their edge set under a rotation or a finite displacement: the criteria are therefore
quantiles over the batch, with the fraction allowed to deviate written in each test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, R = 50, 4096


@pytest.fixture(scope="module")
def setup(weights):
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    mol = synth.synth_molecule(N, seed=N)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=3)).cuda()
    eng = Engine(mol["params"], weights, device=0)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    e, f = eng.energy_force(pos)
    return eng, pos, e.clone(), f.clone()


def _random_rotations(n, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    rot = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                       2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).reshape(n, 3, 3)
    return rot


def test_translation_invariance(setup):
    eng, pos, e0, f0 = setup
    g = torch.Generator().manual_seed(1)
    shift = (torch.rand(R, 1, 3, generator=g) - 0.5).cuda()          # up to 0.5 nm per axis, one shift per conformer
    e1, f1 = eng.energy_force(pos + shift)
    rel = ((e1 - e0).abs() / e0.abs()).cpu()
    # differences of shifted fp32 coordinates round differently: 1e-4 relative for 99.5 % of the conformers
    assert float(torch.quantile(rel, 0.995)) < 1e-4
    frel = ((f1 - f0) ** 2).sum((1, 2)).sqrt() / (f0 ** 2).sum((1, 2)).sqrt()
    assert float(torch.quantile(frel.cpu(), 0.99)) < 1e-3


def test_rotation_invariance_and_force_covariance(setup):
    eng, pos, e0, f0 = setup
    rot = _random_rotations(R, 2).cuda()
    com = pos.double().mean(1, keepdim=True)
    pos_r = (torch.einsum("rij,rnj->rni", rot, pos.double() - com) + com).float()
    e1, f1 = eng.energy_force(pos_r)
    rel = ((e1 - e0).abs() / e0.abs()).cpu()
    # the reference's per-component 1e-6 eps in the pair distance is not rotation invariant (relative 1e-5 on r)
    assert float(torch.quantile(rel, 0.99)) < 1e-4
    f0_r = torch.einsum("rij,rnj->rni", rot, f0.double()).float()
    frel = (((f1 - f0_r) ** 2).sum((1, 2)).sqrt() / (f0 ** 2).sum((1, 2)).sqrt()).cpu()
    # the eps vector (1e-6 per component) turns against the molecule: every directed pair term moves by ~1e-6 / r of
    # its own size, and the net force is a small difference of large pair terms (measured: median 3e-4)
    assert float(torch.quantile(frel, 0.5)) < 1e-3
    assert float(torch.quantile(frel, 0.99)) < 2e-2


def test_net_force_and_torque_vanish(setup):
    eng, pos, e0, f0 = setup
    f = f0.double()
    frms = (f ** 2).sum(2).mean(1).sqrt()                              # per-conformer RMS force per atom
    net = f.sum(1).norm(dim=1) / (frms * np.sqrt(N))
    assert float(net.max()) < 1e-4                                     # exact by construction (pair terms are antisymmetric)
    x = pos.double() - pos.double().mean(1, keepdim=True)
    torque = torch.cross(x, f, dim=2).sum(1).norm(dim=1) / (frms * np.sqrt(N) * x.norm(dim=2).mean(1))
    # the eps-distance breaks rotation invariance at the 1e-5 level, fp32 roundoff in r-bar adds to it
    assert float(torch.quantile(torque.cpu(), 0.99)) < 1e-3


def test_forces_are_the_energy_gradient(setup):
    eng, pos, e0, f0 = setup
    g = torch.Generator().manual_seed(4)
    d = torch.randn(R, N, 3, generator=g).cuda()
    d = d / d.reshape(R, -1).norm(dim=1).reshape(R, 1, 1)              # unit direction per conformer
    h = 2e-4                                                           # nm
    ep, _ = eng.energy_force(pos + h * d)
    em, _ = eng.energy_force(pos - h * d)
    num = (ep.double() - em.double()) / (2 * h)
    ana = -(f0.double() * d.double()).sum((1, 2))
    scale = (f0.double() ** 2).sum((1, 2)).sqrt()                      # |F| of the conformer
    err = ((num - ana).abs() / scale).cpu()
    # fp32 energies (~1e3 kJ/mol, 6e-5 absolute resolution) limit the central difference to ~0.3 / |F|.  A conformer has
    # ~5000 directed pairs per nm of separation around the 0.6 nm cutoff, so about a third of the batch changes its
    # edge set inside +-h (a finite energy step) and falls outside: the criterion is the median
    # and the 60 % quantile
    assert float(torch.quantile(err, 0.5)) < 2e-3
    assert float(torch.quantile(err, 0.6)) < 1e-2


def test_replica_is_independent_of_its_batch(setup):
    eng, pos, e0, f0 = setup
    perm = torch.randperm(R, generator=torch.Generator().manual_seed(5)).cuda()
    e1, f1 = eng.energy_force(pos[perm].contiguous())
    assert torch.equal(e1, e0[perm]) and torch.equal(f1, f0[perm])     # bit-identical: no cross-replica arithmetic
