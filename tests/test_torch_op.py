"""The TORCH_LIBRARY shim (csrc/torch_op.cpp): registration and scripting on CPU, numerics on the GPU."""
import io
import os

import numpy as np
import pytest
import torch

from conftest import rel_rms


def test_operator_library_loads_and_registers():
    from gnnimplicitsolvent_b200 import torch_op
    torch_op.load()
    assert hasattr(torch.ops.gnnis, "energy_force") and hasattr(torch.ops.gnnis, "energy")
    schema = str(torch.ops.gnnis.energy_force.default._schema)
    assert "Tensor pos, int handle, int n_rep, int flags" in schema and "(Tensor, Tensor)" in schema


def test_modules_script_and_serialise_without_gpu():
    from gnnimplicitsolvent_b200 import torch_op
    torch_op.load()
    for cls in (torch_op.ScriptableEnergy, torch_op.ScriptableEnergyForces):
        m = torch.jit.script(cls(1234, 2).eval())
        buf = io.BytesIO()
        torch.jit.save(m, buf)
        buf.seek(0)
        m2 = torch.jit.load(buf)
        assert "gnnis::energy" in str(m2.graph)
        with pytest.raises(Exception):          # no CPU kernel is registered: the call must fail loudly, not fall back
            m2(torch.zeros(4, 3))


@pytest.mark.gpu
def test_scripted_modules_match_engine():
    from gnnimplicitsolvent_b200 import synth, torch_op
    from gnnimplicitsolvent_b200.engine import Engine
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "gnn_weights.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    N, R = 21, 6
    mol = synth.synth_molecule(N, seed=8)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=4)).cuda()
    eng = Engine(mol["params"], sd)
    eng.set_replicas(R, [0, 3, 5, 7, 11, 13], [78.5, 47.24, 32.6, 24.9, 20.5, 10.4])
    e_ref, f_ref = eng.energy_force(pos)

    # (energy, forces) module, through a save / load round trip as TorchForce would consume it
    m = torch_op.script_engine(eng, outputs_forces=True)
    buf = io.BytesIO()
    torch.jit.save(m, buf)
    buf.seek(0)
    m = torch.jit.load(buf)
    flat = pos.reshape(R * N, 3)                        # OpenMM passes [n_particles, 3]
    e, f = m(flat, torch.tensor(-1), torch.tensor(78.5))
    assert abs(float(e) - float(e_ref.double().sum())) <= 1e-6 * abs(float(e_ref.double().sum()))
    np.testing.assert_array_equal(f.reshape(R, N, 3).cpu().numpy(), f_ref.cpu().numpy())

    # energy-only module: backward fills positions.grad with -forces
    m2 = torch_op.script_engine(eng, outputs_forces=False)
    p = flat.clone().requires_grad_(True)
    m2(p).backward()
    np.testing.assert_array_equal((-p.grad).reshape(R, N, 3).cpu().numpy(), f_ref.cpu().numpy())
    # per-replica weights reach the right atoms
    p2 = flat.clone().requires_grad_(True)
    w = torch.arange(1, R + 1, device="cuda", dtype=torch.float32)
    (torch.ops.gnnis.energy(p2, m2.handle, R, 0) * w).sum().backward()
    assert rel_rms((-p2.grad).reshape(R, N, 3).cpu().numpy(), (f_ref * w.view(R, 1, 1)).cpu().numpy()) < 1e-6


@pytest.mark.gpu
def test_create_gnn_model_jit_matches_eager():
    """create_gnn_model(..., jit=True) (helper_functions.py:750-790 with jit) returns a TorchScript module whose
    forward / backward agree with the eager drop-in model."""
    from gnnimplicitsolvent_b200 import GNN_Models, synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "gnn_weights.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    N, R = 13, 4
    mol = synth.synth_molecule(N, seed=2)
    pos = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=9)).reshape(R * N, 3).cuda()
    kw = dict(num_reps=R, solvent_models=[3] * R, solvent_dielectric=[47.24] * R, device="cuda")
    eager = GNN_Models.create_gnn_model(mol["params"], sd, **kw)
    scripted = GNN_Models.create_gnn_model(mol["params"], sd, jit=True, **kw)
    assert isinstance(scripted, torch.jit.ScriptModule)
    p1 = pos.clone().requires_grad_(True)
    e1 = eager(p1, torch.tensor(-1), torch.tensor(78.5))
    e1.backward()
    p2 = pos.clone().requires_grad_(True)
    e2 = scripted(p2, torch.tensor(-1), torch.tensor(78.5))
    e2.backward()
    assert abs(float(e1) - float(e2)) <= 1e-6 * abs(float(e1))
    np.testing.assert_array_equal(p1.grad.cpu().numpy(), p2.grad.cpu().numpy())
