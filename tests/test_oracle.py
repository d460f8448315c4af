"""CPU: the oracle restatement against (a) the committed golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and (b) the imported reference itself when /root/reference exists."""
import numpy as np
import pytest
import torch

from conftest import CASES, load_case, rel_rms
from oracle import gnn_oracle as O
from oracle import ref_harness as H


@pytest.fixture(autouse=True)
def _single_thread():
    """Bit-identity with the golden vectors holds for the sequential (1-thread) CPU summation order the
    fixtures were generated with; multi-threaded index_add / sum reorder fp32 additions (~1e-7)."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name, weights):
    c = load_case(name)
    o = O.evaluate(c["params"], c["pos"], c["solvent_ids"], c["dielectrics"], weights)
    # fp32 restatement follows the reference's op order: energies bit-identical, forces to fp32 noise
    assert float(o["energy_total"]) == float(c["energy_total"])
    assert rel_rms(o["forces"].numpy(), c["forces"]) < 2e-6
    np.testing.assert_array_equal(o["e_atom"].numpy(), c["e_atom"])
    np.testing.assert_array_equal(o["sa_atom"].numpy(), c["sa_atom"])
    np.testing.assert_array_equal(o["born"].numpy(), c["born"])
    od = O.evaluate(c["params"], c["pos"], c["solvent_ids"], c["dielectrics"], weights, delta=True)
    assert float(od["energy_total"]) == float(c["energy_total_delta"])
    assert rel_rms(od["forces"].numpy(), c["forces_delta"]) < 2e-6


@pytest.mark.parametrize("name", CASES)
def test_neighbor_oracle_bit_exact(name):
    """The numpy fp32 predicate (fma-chain eps-distance < 0.6f) reproduces the reference's own edge
    selection exactly: same edges, same order (target, then source), same fp32 distances."""
    c = load_case(name)
    row_ptr, col, r = O.neighbor_list_numpy(c["pos"])
    np.testing.assert_array_equal(col, c["edge_src"])
    np.testing.assert_array_equal(r, c["edge_r"])
    R, N, _ = c["pos"].shape
    tgt = np.repeat(np.arange(R * N), np.diff(row_ptr))
    np.testing.assert_array_equal(tgt, c["edge_tgt"])


def test_single_solvent_override_matches_golden(weights):
    c = load_case("mixed_n21")
    o = O.evaluate(c["params"], c["pos"][:1], c["solvent_ids"][:1], c["dielectrics"][:1], weights)
    assert float(o["energy_total"]) == float(c["energy_rep0_override"])
    assert rel_rms(o["forces"].numpy()[0], c["forces_rep0_override"]) < 2e-6


def test_eps_distance_formula_is_fma_chain():
    """torch's CPU PairwiseDistance == sqrt(fma(dz,dz,fma(dy,dy,dx*dx))) bit for bit (the predicate the
    CUDA neighbour kernel implements with __fmaf_rn)."""
    g = torch.Generator().manual_seed(1)
    a = torch.rand(50000, 3, generator=g) * 1.5
    b = torch.rand(50000, 3, generator=g) * 1.5
    rt = torch.nn.functional.pairwise_distance(a, b).numpy()
    d = (a.numpy() - b.numpy()) + np.float32(1e-6)
    d64 = d.astype(np.float64)
    acc = (d64[:, 0] * d64[:, 0]).astype(np.float32)
    acc = (d64[:, 1] * d64[:, 1] + acc.astype(np.float64)).astype(np.float32)
    acc = (d64[:, 2] * d64[:, 2] + acc.astype(np.float64)).astype(np.float32)
    np.testing.assert_array_equal(np.sqrt(acc), rt)


def test_fp64_oracle_finite_difference_forces(weights):
    """Autograd forces of the fp64 oracle agree with central finite differences (pins the force sign/convention)."""
    c = load_case("c1_n13_dmso")
    pos = c["pos"][:1].astype(np.float64)
    o = O.evaluate(c["params"], pos, c["solvent_ids"][:1], c["dielectrics"][:1], weights, dtype=torch.float64)
    h = 1e-5
    for (a, k) in [(0, 0), (5, 1), (12, 2)]:
        p1, p2 = pos.copy(), pos.copy()
        p1[0, a, k] += h
        p2[0, a, k] -= h
        e1 = float(O.evaluate(c["params"], p1, c["solvent_ids"][:1], c["dielectrics"][:1], weights, dtype=torch.float64, forces=False)["energy_total"])
        e2 = float(O.evaluate(c["params"], p2, c["solvent_ids"][:1], c["dielectrics"][:1], weights, dtype=torch.float64, forces=False)["energy_total"])
        fd = -(e1 - e2) / (2 * h)
        assert abs(fd - float(o["forces"][0, a, k])) < 1e-3 * max(1.0, abs(fd))


@pytest.mark.skipif(H.reference_root() is None, reason="reference tree only exists on the build box")
def test_oracle_vs_imported_reference_random_weights():
    """Live cross-check with seeded random-init weights of the shipped architecture (not only the checkpoint)."""
    from gnnimplicitsolvent_b200 import synth
    sd = O.random_state_dict(seed=3)
    mol = synth.synth_molecule(17, seed=5)
    pos = synth.conformers(mol["pos"], 3, 0.01, seed=7)
    sids, diel = [2, 9, 30], [33.0, 2.38, 46.53]
    model = H.build_reference_run_model(mol["params"].tolist(), sd, 3, sids, diel)
    p = torch.tensor(pos.reshape(-1, 3), requires_grad=True)
    e = model(p, torch.tensor(-1), torch.tensor(78.5))
    e.backward()
    o = O.evaluate(mol["params"], pos, sids, diel, sd)
    assert float(e) == float(o["energy_total"])
    assert rel_rms(o["forces"].numpy(), (-p.grad).reshape(3, 17, 3).numpy()) < 2e-6


def test_activation_stash_codec_arithmetic():
    """The 16-bit fixed-point code of silu'(x) used between the forward and adjoint edge kernels (csrc/fastmath.cuh,
    stash_enc2 / stash_dec2), re-stated in numpy with the same fp32 operations: one FMA whose rounding lands the code in
    the mantissa of a float in [2^23, 2^24), decode by one FMA with exactly representable constants."""
    K, ENC_OFF = np.float32(52428.8), np.float32(8395264.0)
    INV, DEC_OFF = np.float32(1.9073486328125e-05), np.float32(-160.126953125)
    assert float(INV) == 1.25 / 65536 and float(DEC_OFF) == -(2 ** 23 + 6656) * 1.25 / 65536    # exact constants
    d = np.linspace(-0.0998, 1.0998, 200001, dtype=np.float32)       # the range of silu'
    t = (d.astype(np.float64) * np.float64(K) + np.float64(ENC_OFF)).astype(np.float32)          # fma, one rounding
    bits = t.view(np.uint32)
    assert np.all((bits >> 23) == (0x4B000000 >> 23))                                             # exponent of [2^23, 2^24)
    code = bits & 0xFFFF
    assert code.min() >= 1 and code.max() <= 65534                                                # never wraps
    f = (np.uint32(0x4B000000) | code).view(np.float32)                                           # PRMT with 0x4B00
    dec = (f.astype(np.float64) * np.float64(INV) + np.float64(DEC_OFF)).astype(np.float32)
    err = np.abs(dec.astype(np.float64) - d.astype(np.float64))
    assert err.max() <= 0.5 * 1.25 / 65536 * 1.01 + 1e-7, err.max()                              # half a code step: 9.6e-6
    assert abs(float(np.mean(dec.astype(np.float64) - d))) < 2e-7                                 # unbiased
