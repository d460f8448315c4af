"""The drop-in classes of gnnimplicitsolvent_b200.GNN_Models: reference names, constructor arguments, state-dict
contract (CPU) and energy / gradient parity through `forward` + `backward` (GPU)."""
import numpy as np
import pytest
import torch

from conftest import load_case, rel_rms

E_RTOL = 1e-4
F_RRMS = 1e-3

# the 39 entries of MachineLearning/trained_models/GNN.pt["model"]
N_STATE_KEYS = 39


def _run_model(case, weights, cls_name="GNN3_Multisolvent_embedding_run_multiple", **kw):
    from gnnimplicitsolvent_b200 import GNN_Models as M
    R = case["pos"].shape[0]
    model = getattr(M, cls_name)(parameters=case["params"].tolist(), device="cuda" if torch.cuda.is_available() else "cpu",
                                 num_reps=R, solvent_models=[int(s) for s in case["solvent_ids"]],
                                 solvent_dielectric=[float(d) for d in case["dielectrics"]], num_solvents=42, **kw)
    model.load_state_dict(weights)
    return model


def test_state_dict_contract_matches_checkpoint(weights):
    """load_state_dict(GNN.pt['model']) with strict=True: all 39 keys matched, same shapes and dtypes."""
    from gnnimplicitsolvent_b200 import GNN_Models as M
    c = load_case("c1_n13_dmso")
    model = M.GNN3_Multisolvent_embedding_run_multiple(parameters=c["params"].tolist(), device="cpu", num_solvents=42)
    sd = model.state_dict()
    assert len(weights) == N_STATE_KEYS
    assert set(sd.keys()) == set(weights.keys())
    for k, v in weights.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        assert sd[k].dtype == v.dtype, k
    res = model.load_state_dict(weights, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the resampled neck tables built by the constructor equal the checkpoint buffers (SURVEY.md a5 [probe])
    fresh = M.GNN3_Multisolvent_embedding_run_multiple(parameters=c["params"].tolist(), device="cpu", num_solvents=42)
    np.testing.assert_allclose(fresh.state_dict()["aggregate_information._d0"].numpy(),
                               weights["aggregate_information._d0"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(fresh.state_dict()["aggregate_information._m0"].numpy(),
                               weights["aggregate_information._m0"].numpy(), rtol=1e-6)


def test_reference_constructor_signature():
    """Positional / keyword arguments and defaults of the reference constructors are kept."""
    import inspect
    from gnnimplicitsolvent_b200 import GNN_Models as M
    sig = inspect.signature(M.GNN3_Multisolvent_embedding_run_multiple.__init__)
    names = list(sig.parameters)[1:]
    assert names == ["fraction", "radius", "max_num_neighbors", "parameters", "device", "jittable", "num_reps",
                     "gbneck_radius", "unique_radii", "hidden", "solvent_dielectric", "num_solvents",
                     "solvent_models", "hidden_token", "scaling_factor"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["fraction"], d["radius"], d["max_num_neighbors"], d["num_reps"], d["gbneck_radius"]) == (0.1, 0.6, 10000, 1, 10.0)
    assert (d["hidden"], d["num_solvents"], d["hidden_token"], d["scaling_factor"]) == (64, 1, 128, 2.0)
    sig = inspect.signature(M.GNN3_Multisolvent_embedding.__init__)
    assert list(sig.parameters)[1:] == ["fraction", "radius", "max_num_neighbors", "parameters", "device", "jittable",
                                        "unique_radii", "hidden", "num_solvents", "dropout_rate", "hidden_token",
                                        "scaling_factor", "no_batch"]
    assert M.DEFAULT_UNIQUE_RADII == [0.14, 0.117, 0.155, 0.15, 0.21, 0.185, 0.18, 0.17, 0.12, 0.13]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_forward_without_gpu_fails_loudly(weights):
    """There is no CPU fallback: forward raises instead of silently computing elsewhere."""
    from gnnimplicitsolvent_b200.engine import GnnisError
    c = load_case("c1_n13_dmso")
    model = _run_model(c, weights)
    with pytest.raises(GnnisError):
        model(torch.from_numpy(c["pos"].reshape(-1, 3)), -1, 78.5)


def test_set_num_reps_argument_checks(weights):
    c = load_case("c1_n13_dmso")
    model = _run_model(c, weights)
    assert model.set_num_reps(2, [3, 4], [47.24, 20.0]) == 0
    with pytest.raises(IndexError):
        model.set_num_reps(3, [3, 4], [47.24, 20.0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_n13_dmso", "mixed_n21", "c2_n50_water"])
def test_run_model_forward_backward_vs_golden(name, weights):
    """TorchForce contract: scalar energy, energy.backward() fills positions.grad = -forces."""
    c = load_case(name)
    model = _run_model(c, weights)
    pos = torch.from_numpy(c["pos"].reshape(-1, 3)).cuda().requires_grad_(True)
    e = model(pos, torch.tensor(-1), torch.tensor(78.5))
    assert e.dim() == 0
    e.backward()
    assert abs(float(e) - float(c["energy_total"])) <= E_RTOL * abs(float(c["energy_total"]))
    assert rel_rms(-pos.grad.cpu().numpy().reshape(c["forces"].shape), c["forces"]) < F_RRMS
    # scaled upstream gradient
    pos2 = pos.detach().clone().requires_grad_(True)
    (2.5 * model(pos2, -1, 78.5)).backward()
    assert torch.allclose(pos2.grad, 2.5 * pos.grad, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_delta_and_split_models_vs_golden(weights):
    c = load_case("mixed_n21")
    delta = _run_model(c, weights, "GNN3_Multisolvent_embedding_run_multiple_Delta")
    pos = torch.from_numpy(c["pos"].reshape(-1, 3)).cuda().requires_grad_(True)
    e = delta(pos)
    e.backward()
    assert abs(float(e) - float(c["energy_total_delta"])) <= E_RTOL * abs(float(c["energy_total_delta"]))
    assert rel_rms(-pos.grad.cpu().numpy().reshape(c["forces_delta"].shape), c["forces_delta"]) < F_RRMS
    split = _run_model(c, weights, "GNN3_Multisolvent_embedding_run_multiple_split")
    en, sa = split(pos.detach(), -1, 78.5)
    assert en.shape == (pos.shape[0], 1) and sa.shape == (pos.shape[0], 1)
    np.testing.assert_allclose(en.cpu().numpy().ravel(), c["e_atom"].ravel(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(sa.cpu().numpy().ravel(), c["sa_atom"].ravel(), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_single_solvent_override_through_forward(weights):
    """solvent_model != -1 evaluates in that solvent (GNN_Models.py:780-786)."""
    c = load_case("mixed_n21")
    from gnnimplicitsolvent_b200 import GNN_Models as M
    model = M.GNN3_Multisolvent_embedding_run_multiple(parameters=c["params"].tolist(), device="cuda", num_solvents=42)
    model.load_state_dict(weights)
    pos = torch.from_numpy(c["pos"][:1].reshape(-1, 3)).cuda().requires_grad_(True)
    e = model(pos, torch.tensor(int(c["solvent_ids"][0])), torch.tensor(float(c["dielectrics"][0])))
    e.backward()
    assert abs(float(e) - float(c["energy_rep0_override"])) <= E_RTOL * abs(float(c["energy_rep0_override"]))
    assert rel_rms(-pos.grad.cpu().numpy().reshape(-1, 3), c["forces_rep0_override"]) < F_RRMS


@pytest.mark.gpu
def test_data_model_vs_oracle(weights):
    """forward(data) -> (energy, forces) with the torch_cluster radius-graph predicate."""
    from types import SimpleNamespace
    from gnnimplicitsolvent_b200 import GNN_Models as M
    from oracle import gnn_oracle as O
    c = load_case("mixed_n21")
    R, N, _ = c["pos"].shape
    o = O.evaluate(c["params"], c["pos"], c["solvent_ids"], c["dielectrics"], weights, predicate="data")
    data = SimpleNamespace(
        pos=torch.from_numpy(c["pos"].reshape(-1, 3)).cuda(),
        atom_features=torch.tensor(c["params"], dtype=torch.float32).repeat(R, 1),
        batch=torch.arange(R).repeat_interleave(N),
        solvent_model_tensor=torch.tensor(c["solvent_ids"]).repeat_interleave(N),
        solvent_dielectric=torch.tensor(c["dielectrics"], dtype=torch.float32).repeat_interleave(N))
    model = M.GNN3_Multisolvent_embedding(parameters=None, device="cuda", num_solvents=42, no_batch=False)
    model.load_state_dict(weights)
    energy, forces = model(data)
    assert energy.shape == (R, 1) and forces.shape == (R * N, 3)
    np.testing.assert_allclose(energy.cpu().numpy().ravel(), o["energy_rep"].numpy(), rtol=E_RTOL)
    assert rel_rms(forces.cpu().numpy().reshape(R, N, 3), o["forces"].numpy()) < F_RRMS
    model_nb = M.GNN3_Multisolvent_embedding(parameters=None, device="cuda", num_solvents=42)
    model_nb.load_state_dict(weights)
    e1, _ = model_nb(data)
    assert e1.shape == (1, 1)
    assert abs(float(e1) - float(o["energy_total"])) <= E_RTOL * abs(float(o["energy_total"]))


def test_gbneck2_parameter_table_layout():
    """N2: [q, rho, s, alpha, beta, gamma, radindex] with rho = r - 0.0195141, s = screen * rho and the UNSORTED
    radindex of the default unique-radii list (GNN_Trainer.py:861-867; carbon 0.17 -> index 7)."""
    from gnnimplicitsolvent_b200 import gbneck, synth
    mol = synth.synth_molecule(21, seed=3)
    els = ["H" if e == "HN" else e for e in mol["elements"]]
    p, radii = gbneck.gbneck2_parameters(els, mol["bonds"], mol["params"][:, 0])
    np.testing.assert_allclose(p, mol["params"], rtol=0, atol=1e-12)        # the synthetic generator follows the same rules
    c = els.index("C")
    assert p[c, 6] == 7 and abs(p[c, 1] - (0.17 - 0.0195141)) < 1e-12 and abs(p[c, 2] - 1.058554 * p[c, 1]) < 1e-12
    assert radii == gbneck.DEFAULT_UNIQUE_RADII


def test_gbneck2_element_table_vs_reference_notebook_output():
    """N2 pinned where the reference allows it: OpenMM is absent, but the reference's own notebook prints the first
    columns of `atomfeatures` rows that its Trainer.get_gbneck2_param produced (/root/reference/Data/deposit_database.ipynb,
    cell 77 output: `[[q, rho, s, alpha, ...`).  rho = r - 0.0195141, s = screen * rho and alpha of C, O, N and S from
    gbneck.gbneck2_parameters must reproduce those printed digits."""
    from gnnimplicitsolvent_b200 import gbneck
    printed = {"C": (0.1504859, 0.15929745, 0.733756), "O": (0.1304859, 0.13845063, 0.867814),
               "N": (0.1354859, 0.09939232, 0.503364), "S": (0.1604859, -0.11289686, 0.867814)}
    els = list(printed)
    p, _ = gbneck.gbneck2_parameters(els, [], [0.0] * len(els))
    for i, el in enumerate(els):
        rho, s_, alpha = printed[el]
        assert abs(p[i, 1] - rho) < 5e-8 and abs(p[i, 2] - s_) < 5e-8 and abs(p[i, 3] - alpha) < 5e-7, (el, p[i])


@pytest.mark.gpu
def test_data_model_ragged_batch(weights):
    """A PyG-style batch mixing two different molecules (13 and 21 atoms, interleaved graphs, mixed solvents)."""
    from types import SimpleNamespace
    from gnnimplicitsolvent_b200 import GNN_Models as M
    from oracle import gnn_oracle as O
    ca, cb = load_case("c1_n13_dmso"), load_case("mixed_n21")
    order = [("a", 0), ("b", 0), ("b", 1), ("a", 1), ("b", 2)]
    pos, feats, batch, sid, die, ref_e, ref_f = [], [], [], [], [], [], []
    oa = O.evaluate(ca["params"], ca["pos"][:2], ca["solvent_ids"][:2], ca["dielectrics"][:2], weights, predicate="data")
    ob = O.evaluate(cb["params"], cb["pos"][:3], cb["solvent_ids"][:3], cb["dielectrics"][:3], weights, predicate="data")
    for g, (which, k) in enumerate(order):
        c, o = (ca, oa) if which == "a" else (cb, ob)
        N = c["pos"].shape[1]
        pos.append(c["pos"][k]), feats.append(np.asarray(c["params"], dtype=np.float32))
        batch += [g] * N
        sid += [int(c["solvent_ids"][k])] * N
        die += [float(c["dielectrics"][k])] * N
        ref_e.append(float(o["energy_rep"][k])), ref_f.append(o["forces"][k].numpy())
    data = SimpleNamespace(pos=torch.from_numpy(np.concatenate(pos)).cuda(), atom_features=torch.from_numpy(np.concatenate(feats)),
                           batch=torch.tensor(batch), solvent_model_tensor=torch.tensor(sid),
                           solvent_dielectric=torch.tensor(die, dtype=torch.float32))
    model = M.GNN3_Multisolvent_embedding(parameters=None, device="cuda", num_solvents=42, no_batch=False)
    model.load_state_dict(weights)
    energy, forces = model(data)
    assert energy.shape == (5, 1) and forces.shape == (len(batch), 3)
    np.testing.assert_allclose(energy.cpu().numpy().ravel(), np.array(ref_e), rtol=E_RTOL)
    assert rel_rms(forces.cpu().numpy(), np.concatenate(ref_f)) < F_RRMS
