"""GPU parity tests (run under gpurun on a B200): the CUDA path through the C ABI vs the golden vectors
produced by the unmodified reference and vs the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): neighbour lists bit-exact; energies 1e-4 relative;
forces 1e-3 relative RMS (fp32 mode)."""
import numpy as np
import pytest
import torch

from conftest import CASES, load_case, rel_rms

pytestmark = pytest.mark.gpu

E_RTOL = 1e-4
F_RRMS = 1e-3


def _engine(case, weights, **kw):
    from gnnimplicitsolvent_b200.engine import Engine
    eng = Engine(case["params"], weights, device=0, **kw)
    R = case["pos"].shape[0]
    eng.set_replicas(R, case["solvent_ids"], case["dielectrics"])
    return eng


@pytest.mark.parametrize("name", CASES)
def test_neighbor_list_bit_exact(name, weights):
    c = load_case(name)
    eng = _engine(c, weights)
    row_ptr, col, r = eng.build_neighbors(torch.from_numpy(c["pos"]).cuda())
    R, N, _ = c["pos"].shape
    deg = np.diff(row_ptr.cpu().numpy())
    tgt = np.repeat(np.arange(R * N), deg)
    np.testing.assert_array_equal(col.cpu().numpy(), c["edge_src"])
    np.testing.assert_array_equal(tgt, c["edge_tgt"])
    np.testing.assert_array_equal(r.cpu().numpy(), c["edge_r"])     # fp32 distances bit-identical


@pytest.mark.parametrize("name", CASES)
def test_energy_forces_vs_golden(name, weights):
    c = load_case(name)
    eng = _engine(c, weights)
    pos = torch.from_numpy(c["pos"]).cuda()
    e, f = eng.energy_force(pos)
    torch.cuda.synchronize()
    assert eng.last_edge_count() == len(c["edge_src"])
    born, _ = eng.born_radii()
    np.testing.assert_allclose(born.cpu().numpy(), c["born"], rtol=2e-5)
    e_atom, sa_atom = eng.atom_energies()
    np.testing.assert_allclose(sa_atom.cpu().numpy(), c["sa_atom"], rtol=1e-4, atol=1e-5)
    R, N, _ = c["pos"].shape
    e_rep_ref = (c["e_atom"] + c["sa_atom"]).reshape(R, N).astype(np.float64).sum(1)
    np.testing.assert_allclose(e.cpu().numpy(), e_rep_ref, rtol=E_RTOL)
    assert abs(float(e.double().sum()) - float(c["energy_total"])) <= E_RTOL * abs(float(c["energy_total"]))
    assert rel_rms(f.cpu().numpy(), c["forces"]) < F_RRMS


@pytest.mark.parametrize("name", ["mixed_n21", "c2_n50_water", "n100_mixed"])
def test_delta_mode_vs_golden(name, weights):
    c = load_case(name)
    eng = _engine(c, weights)
    e, f = eng.energy_force(torch.from_numpy(c["pos"]).cuda(), delta=True)
    assert abs(float(e.double().sum()) - float(c["energy_total_delta"])) <= E_RTOL * abs(float(c["energy_total_delta"]))
    assert rel_rms(f.cpu().numpy(), c["forces_delta"]) < F_RRMS


def test_single_replica_override(weights):
    """solvent_model != -1 path (GNN_Models.py:780-786): one replica, one solvent."""
    c = load_case("mixed_n21")
    from gnnimplicitsolvent_b200.engine import Engine
    eng = Engine(c["params"], weights)
    eng.set_replicas(1, c["solvent_ids"][:1], c["dielectrics"][:1])
    e, f = eng.energy_force(torch.from_numpy(c["pos"][:1]).cuda())
    assert abs(float(e[0]) - float(c["energy_rep0_override"])) <= E_RTOL * abs(float(c["energy_rep0_override"]))
    assert rel_rms(f.cpu().numpy()[0], c["forces_rep0_override"]) < F_RRMS


@pytest.mark.parametrize("N,R", [(7, 3), (13, 37), (33, 9), (64, 5), (65, 3), (130, 2)])
def test_vs_oracle_random_weights(N, R):
    """Seeded random-init weights, ragged unit packing (R not a multiple of the unit size), both table modes
    (N <= 64 shared-memory tables, N > 64 global tables)."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    sd = O.random_state_dict(seed=N)
    mol = synth.synth_molecule(N, seed=100 + N)
    pos = synth.conformers(mol["pos"], R, 0.01, seed=N * 7 + R)
    sids = [(3 * r) % 39 for r in range(R)]
    diel = [synth.SOLVENT_DIELECTRICS[s] for s in sids]
    o = O.evaluate(mol["params"], pos, sids, diel, sd)
    eng = Engine(mol["params"], sd)
    eng.set_replicas(R, sids, diel)
    e, f = eng.energy_force(torch.from_numpy(pos).cuda())
    np.testing.assert_allclose(e.cpu().numpy(), o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
    assert rel_rms(f.cpu().numpy(), o["forces"].numpy()) < F_RRMS
    row_ptr, col, r = eng.build_neighbors(torch.from_numpy(pos).cuda())
    rp, cl, rr = O.neighbor_list_numpy(pos)
    np.testing.assert_array_equal(row_ptr.cpu().numpy(), rp)
    np.testing.assert_array_equal(col.cpu().numpy(), cl)


def test_gb_only_mode_vs_oracle():
    """Plain GB-Neck2 (no GNN): validates the hand-written Born/GB adjoint on its own."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    mol = synth.synth_molecule(30, seed=3)
    R = 4
    pos = synth.conformers(mol["pos"], R, 0.01, seed=5)
    x = torch.tensor(mol["params"], dtype=torch.float32).repeat(R, 1)
    p = torch.tensor(pos.reshape(-1, 3), requires_grad=True)
    ei = O.all_pairs_index(30, R)
    r = O.eps_distance(p, ei[0], ei[1])
    d0, m0 = O.neck_tables()
    B, _ = O.born_radii(x, ei[0], ei[1], r, d0, m0)
    diel = torch.full((R * 30, 1), 78.5)
    e = O.gb_energy(torch.stack((B, x[:, 0]), 1), ei[0], ei[1], r, diel)
    e.sum().backward()
    eng = Engine(mol["params"], None)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    eg, fg = eng.energy_force(torch.from_numpy(pos).cuda(), gb_only=True)
    np.testing.assert_allclose(eg.cpu().numpy(), e.detach().reshape(R, 30).sum(1).numpy(), rtol=E_RTOL)
    assert rel_rms(fg.cpu().numpy(), (-p.grad).reshape(R, 30, 3).numpy()) < F_RRMS


def test_data_path_predicate_bit_exact():
    """torch_cluster.radius_graph semantics (sum of squares < r^2), SURVEY.md §8 a4'."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    mol = synth.synth_molecule(40, seed=11)
    pos = synth.conformers(mol["pos"], 6, 0.02, seed=2)
    eng = Engine(mol["params"], None)
    eng.set_replicas(6, [0] * 6, [78.5] * 6)
    row_ptr, col, _ = eng.build_neighbors(torch.from_numpy(pos).cuda(), predicate="data")
    rp, cl, _ = O.neighbor_list_numpy(pos, predicate="data")
    np.testing.assert_array_equal(row_ptr.cpu().numpy(), rp)
    np.testing.assert_array_equal(col.cpu().numpy(), cl)


def test_deterministic_bitwise(weights):
    """No float atomics anywhere: two evaluations of the same input are bit-identical."""
    c = load_case("mixed_n50")
    eng = _engine(c, weights)
    pos = torch.from_numpy(c["pos"]).cuda()
    e1, f1 = eng.energy_force(pos)
    e2, f2 = eng.energy_force(pos)
    assert torch.equal(e1, e2) and torch.equal(f1, f2)


def test_replica_invariance_full_size(weights):
    """Size-independent property at a large batch: every replica of an identical-conformer batch gives the
    same energy/forces as a single replica (replicas are independent)."""
    c = load_case("c2_n50_water")
    R = 1024
    pos1 = c["pos"][:1]
    from gnnimplicitsolvent_b200.engine import Engine
    eng = Engine(c["params"], weights)
    eng.set_replicas(R, [0] * R, [78.5] * R)
    pos = torch.from_numpy(np.repeat(pos1, R, axis=0)).cuda()
    e, f = eng.energy_force(pos)
    # 1024 replicas on 296 warp groups = three full waves of whole replicas + 136 replicas that are split in two (tail
    # split, Ctx::split_from): bit-equal inside either set, equal up to the summation order of the split between them
    T = R - R % 296
    assert "replicas %d.." % T in eng.describe()
    assert torch.equal(e[:T], e[:1].expand(T)) and torch.equal(f[:T], f[:1].expand(T, -1, -1))
    assert torch.equal(e[T:], e[T:T + 1].expand(R - T)) and torch.equal(f[T:], f[T:T + 1].expand(R - T, -1, -1))
    assert abs(float(e[T]) - float(e[0])) <= 2e-6 * abs(float(e[0])) and rel_rms(f[T].cpu().numpy(), f[0].cpu().numpy()) < 2e-5
    assert abs(float(e[0]) - float((c["e_atom"] + c["sa_atom"])[:50].sum())) < E_RTOL * abs(float(e[0]))


def test_host_api_matches_device_api(weights):
    c = load_case("mixed_n21")
    eng = _engine(c, weights)
    pos = torch.from_numpy(c["pos"]).pin_memory()
    e_h = torch.empty(c["pos"].shape[0]).pin_memory()
    f_h = torch.empty(c["pos"].shape).pin_memory()
    eng.energy_force_host(pos, e_h, f_h)
    e, f = eng.energy_force(pos.cuda())
    assert torch.equal(e.cpu(), e_h) and torch.equal(f.cpu(), f_h)


def test_graph_replay_and_invalidation(weights, monkeypatch):
    """Evaluations with recurring buffers are replayed from a CUDA graph inside the library: results must be bit-equal
    to plain stream launches (GNNIS_GRAPHS=0), more buffer sets than cache entries must work, and every setter must
    invalidate what was captured (here: other solvents / dielectrics, then another replica count)."""
    from gnnimplicitsolvent_b200.engine import Engine
    c = load_case("mixed_n21")
    R = c["pos"].shape[0]
    monkeypatch.setenv("GNNIS_GRAPHS", "0")
    plain = Engine(c["params"], weights, device=0)
    monkeypatch.setenv("GNNIS_GRAPHS", "1")
    eng = Engine(c["params"], weights, device=0)
    for e_ in (plain, eng):
        e_.set_replicas(R, c["solvent_ids"], c["dielectrics"])
    g = torch.Generator().manual_seed(11)
    sets = [(torch.from_numpy(c["pos"]) + 0.001 * k * torch.randn(c["pos"].shape, generator=g)).float().cuda() for k in range(11)]
    outs = [(torch.empty(R, device="cuda"), torch.empty(c["pos"].shape, device="cuda")) for _ in sets]
    for rnd in range(4):                       # round 0: first sighting, round 1: capture, rounds 2-3: replay
        for k, p in enumerate(sets):           # 11 buffer sets > 8 cache entries: least-recently-used replacement
            eng.energy_force(p, out_energy=outs[k][0], out_force=outs[k][1])
            e_ref, f_ref = plain.energy_force(p)
            assert torch.equal(outs[k][0], e_ref) and torch.equal(outs[k][1], f_ref), (rnd, k)
    for rnd in range(3):                       # three buffer sets only: these are replayed
        for k in range(3):
            outs[k][0].zero_()
            eng.energy_force(sets[k], out_energy=outs[k][0], out_force=outs[k][1])
            e_ref, f_ref = plain.energy_force(sets[k])
            assert torch.equal(outs[k][0], e_ref) and torch.equal(outs[k][1], f_ref)
    # same buffers, new per-replica solvents: the captured graph must not be reused
    sid2 = [(int(x) + 5) % 39 for x in c["solvent_ids"]]
    die2 = [float(x) * 0.5 + 2.0 for x in c["dielectrics"]]
    for e_ in (plain, eng):
        e_.set_replicas(R, sid2, die2)
    for rnd in range(3):
        eng.energy_force(sets[0], out_energy=outs[0][0], out_force=outs[0][1])
        e_ref, f_ref = plain.energy_force(sets[0])
        assert torch.equal(outs[0][0], e_ref) and torch.equal(outs[0][1], f_ref)
    # fewer replicas through the same buffers
    R2 = R - 1
    for e_ in (plain, eng):
        e_.set_replicas(R2, sid2[:R2], die2[:R2])
    p2 = sets[0][:R2].contiguous()
    for rnd in range(3):
        e2, f2 = torch.empty(R2, device="cuda"), torch.empty(R2, *c["pos"].shape[1:], device="cuda")
        eng.energy_force(p2, out_energy=e2, out_force=f2)
        e_ref, f_ref = plain.energy_force(p2)
        assert torch.equal(e2, e_ref) and torch.equal(f2, f_ref)


def test_pipelined_host_api(weights):
    """submit / wait: results of a stream of batches equal the blocking host call bit for bit, arrive in submission
    order, a third submission in flight is refused (-4) and a wait with nothing in flight is refused."""
    from gnnimplicitsolvent_b200.engine import GnnisError
    c = load_case("mixed_n21")
    eng = _engine(c, weights)
    R = c["pos"].shape[0]
    g = torch.Generator().manual_seed(5)
    batches = [torch.from_numpy(c["pos"]) + 0.002 * k * torch.randn(c["pos"].shape, generator=g) for k in range(5)]
    batches = [b.float().contiguous().pin_memory() for b in batches]
    ref = []
    for b in batches:
        e_h, f_h = torch.empty(R).pin_memory(), torch.empty(c["pos"].shape).pin_memory()
        eng.energy_force_host(b, e_h, f_h)
        ref.append((e_h.clone(), f_h.clone()))
    outs = [(torch.empty(R).pin_memory(), torch.empty(c["pos"].shape).pin_memory()) for _ in batches]
    with pytest.raises(GnnisError):
        eng.energy_force_host_wait()
    for k, b in enumerate(batches):
        eng.energy_force_host_submit(b, outs[k][0], outs[k][1])
        if k == 1:
            with pytest.raises(GnnisError):
                eng.energy_force_host_submit(b, outs[k][0], outs[k][1])   # two already in flight
        if k >= 1:
            eng.energy_force_host_wait()
            assert torch.equal(outs[k - 1][0], ref[k - 1][0]) and torch.equal(outs[k - 1][1], ref[k - 1][1])
    eng.energy_force_host_wait()
    assert torch.equal(outs[-1][0], ref[-1][0]) and torch.equal(outs[-1][1], ref[-1][1])
    # energies only
    e_only = torch.empty(R).pin_memory()
    eng.energy_force_host_submit(batches[0], e_only, None)
    eng.energy_force_host_wait()
    assert torch.equal(e_only, ref[0][0])


@pytest.mark.parametrize("K,N", [(32, 64), (64, 64), (64, 32)])
def test_umma_building_blocks(K, N):
    """tcgen05 building blocks (TMEM-resident A, 128B-swizzled smem B, 3xTF32): D = A B^T to fp32 accuracy."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    mol = synth.synth_molecule(7, seed=1)
    eng = Engine(mol["params"], None)
    g = torch.Generator().manual_seed(K * 100 + N)
    a = torch.randn(128, K, generator=g) * 3.0
    b = torch.randn(N, K, generator=g)
    d = eng.umma_selftest(a, b).cpu().double()
    ref = a.double() @ b.double().T
    err = float((d - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err           # plain TF32 would be ~1e-3


@pytest.mark.parametrize("name", ["mixed_n21", "c2_n50_water", "n100_mixed"])
def test_cuda_core_path_vs_golden(name, weights):
    """The fp32 CUDA-core edge kernels (GNNIS_FLAG_MLP_CUDA_CORES) stay parity-green next to the tcgen05 default."""
    c = load_case(name)
    eng = _engine(c, weights)
    e, f = eng.energy_force(torch.from_numpy(c["pos"]).cuda(), cuda_cores=True)
    assert abs(float(e.double().sum()) - float(c["energy_total"])) <= E_RTOL * abs(float(c["energy_total"]))
    assert rel_rms(f.cpu().numpy(), c["forces"]) < F_RRMS
    e2, f2 = eng.energy_force(torch.from_numpy(c["pos"]).cuda())
    # tensor-core (3xTF32) and CUDA-core paths: energies agree to fp32 rounding; the default adjoint reads silu'(u),
    # silu'(w) back from 16-bit codes (absolute error <= 9.6e-6 per activation derivative), which shows up as a few
    # 1e-5 relative in the forces -- two decades inside the 1e-3 gate
    assert rel_rms(f2.cpu().numpy(), f.cpu().numpy()) < 1e-4
    # (the tensor-core forward takes its sigmoid from MUFU.TANH, the CUDA-core kernels from ex2 + rcp: a few 1e-6)
    np.testing.assert_allclose(e2.cpu().numpy(), e.cpu().numpy(), rtol=2e-5)


def test_cuda_graph_capture(weights):
    """gnnis_energy_force does not synchronise the host: it can be captured into a CUDA graph and replayed."""
    c = load_case("mixed_n50")
    eng = _engine(c, weights)
    pos = torch.from_numpy(c["pos"]).cuda()
    e_ref, f_ref = eng.energy_force(pos)
    e = torch.empty_like(e_ref)
    f = torch.empty_like(f_ref)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.energy_force(pos, out_energy=e, out_force=f)      # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.energy_force(pos, out_energy=e, out_force=f)
    e.zero_()
    f.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(e, e_ref) and torch.equal(f, f_ref)
    # new positions through the same graph
    pos.add_(0.001)
    graph.replay()
    torch.cuda.synchronize()
    e2, f2 = eng.energy_force(pos)
    assert torch.equal(e, e2) and torch.equal(f, f2)


@pytest.mark.parametrize("name", ["c2_n50_water", "n100_mixed"])
def test_tf32x1_mode(name, weights):
    """Opt-in reduced-precision mode of the edge MLPs (GNNIS_FLAG_MLP_TF32X1: one TF32 product per contraction).
    north_star asks for the tolerance of a reduced-precision MLP mode to be stated separately: it misses the fp32
    gates (1e-4 / 1e-3) and is held to energies 2e-3, forces 2e-2 relative; the default path is unaffected."""
    c = load_case(name)
    eng = _engine(c, weights)
    pos = torch.from_numpy(c["pos"]).cuda()
    e3, f3 = eng.energy_force(pos)
    e1, f1 = eng.energy_force(pos, tf32x1=True)
    e_rel = float((e1 - e3).abs().max() / e3.abs().max())
    f_rel = rel_rms(f1.cpu().numpy(), f3.cpu().numpy())
    print(f"tf32x1 vs 3xTF32 [{name}]: energy max rel {e_rel:.2e}, force rel-RMS {f_rel:.2e}")
    assert 0.0 < e_rel < 2e-3 and 0.0 < f_rel < 2e-2          # different from, and close to, the fp32-accurate result
    e3b, f3b = eng.energy_force(pos)                           # the default mode is bit-reproducible afterwards
    assert torch.equal(e3, e3b) and torch.equal(f3, f3b)


# measured on B200 against the reference's goldens (worst of the five cases; printed by the test, quoted in DESIGN.md 5.5)
HALF_MODE_BOUNDS = {"bf16": (5e-3, 8e-2), "fp16": (2e-3, 2e-2)}


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("name", ["c1_n13_dmso", "c2_n50_water", "mixed_n21", "mixed_n50", "n100_mixed"])
def test_half_precision_modes(name, mode, weights):
    """The half-precision MLP modes (GNNIS_FLAG_MLP_BF16, the north star's "bf16 MLP mode", and GNNIS_FLAG_MLP_FP16):
    per-edge contractions as tcgen05.mma.kind::f16 on 16-bit operands (fp32 accumulate), SiLU from MUFU.TANH.
    north_star asks for the force tolerance of this mode to be stated separately from the fp32 one: against the
    REFERENCE's goldens the modes are held to HALF_MODE_BOUNDS (energy max rel, force rel-RMS); they must differ from
    the fp32-accurate default, which stays bit-reproducible afterwards, and be deterministic themselves."""
    c = load_case(name)
    eng = _engine(c, weights)
    pos = torch.from_numpy(c["pos"]).cuda()
    kw = {mode: True}
    e3, f3 = eng.energy_force(pos)
    eb, fb = eng.energy_force(pos, **kw)
    assert bool(torch.isfinite(eb).all() and torch.isfinite(fb).all())
    R, N, _ = c["pos"].shape
    e_ref = (c["e_atom"] + c["sa_atom"]).reshape(R, N).astype(np.float64).sum(1)     # the reference's per-replica energies
    e_rel = float(np.abs(eb.double().cpu().numpy() - e_ref).max() / np.abs(e_ref).max())
    f_rel = rel_rms(fb.cpu().numpy(), c["forces"])
    print(f"{mode} mode vs reference goldens [{name}]: energy max rel {e_rel:.2e}, force rel-RMS {f_rel:.2e}")
    e_max, f_max = HALF_MODE_BOUNDS[mode]
    assert 0.0 < e_rel < e_max and 0.0 < f_rel < f_max
    assert not torch.equal(eb, e3)
    e3b, f3b = eng.energy_force(pos)
    assert torch.equal(e3, e3b) and torch.equal(f3, f3b)
    eb2, fb2 = eng.energy_force(pos, **kw)
    assert torch.equal(eb, eb2) and torch.equal(fb, fb2)


def test_capacity_limited_stash(weights, monkeypatch):
    """A stash that holds less than the worst-case edge capacity (Ctx::stash_cap; here forced with GNNIS_STASH_GB): the
    kernels compare the edge count of the evaluation with it ON THE DEVICE.  A list that fits takes the stash adjoint
    (bit-identical to the unlimited stash), one that does not fit is not stashed and the recomputing adjoint runs
    (bit-identical to GNNIS_STASH=0); below half of the capacity no stash is allocated at all."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine

    def engines(make):
        eng_s = make()
        monkeypatch.setenv("GNNIS_STASH", "0")
        eng_r = make()
        monkeypatch.delenv("GNNIS_STASH")
        return eng_s, eng_r

    def limited(make, hold):
        # ensure_workspace: hold = floor(budget / 768) - 128, rounded down to whole tiles of 128 edges
        monkeypatch.setenv("GNNIS_STASH_GB", repr(((hold + 128) * 768 + 64) / 1e9))
        eng = make()
        monkeypatch.delenv("GNNIS_STASH_GB")
        return eng

    mol = synth.synth_molecule(50, seed=50)
    R = 600                                                    # two full waves of units + a split tail
    pos_big = torch.from_numpy(synth.conformers(mol["pos"], R, 0.01, seed=3)).cuda()

    def make_big():
        eng = Engine(mol["params"], weights, device=0)
        eng.set_replicas(R, [0] * R, [78.5] * R)
        return eng

    cases = [(lambda c=load_case(n): _engine(c, weights), torch.from_numpy(load_case(n)["pos"]).cuda())
             for n in ("c2_n50_water", "n100_mixed")] + [(make_big, pos_big)]
    seen_fit = seen_nofit = 0
    for make, pos in cases:
        Rr, N, _ = pos.shape
        cap = Rr * N * (N - 1)
        eng_s, eng_r = engines(make)
        es, fs = eng_s.energy_force(pos)
        er, fr = eng_r.energy_force(pos)
        E = eng_s.last_edge_count()
        half = -(-cap // 2)
        # (a) holds the list, not the capacity
        hold = max(-(-E // 128) * 128, -(-half // 128) * 128) + 128
        if hold < cap:
            eng = limited(make, hold)
            assert f"adjoint=stash up to {hold} edges" in eng.describe(), eng.describe()
            for _ in range(3):                                 # plain launches, then graph capture and replay
                e, f = eng.energy_force(pos)
                assert torch.equal(e, es) and torch.equal(f, fs)
            seen_fit += 1
        # (b) at least half of the capacity, but not the list: decided on the device, nothing is stashed
        hold = -(-half // 128) * 128
        if hold + 128 <= E:
            eng = limited(make, hold)
            assert f"adjoint=stash up to {hold} edges" in eng.describe(), eng.describe()
            for _ in range(3):
                e, f = eng.energy_force(pos)
                assert torch.equal(e, er) and torch.equal(f, fr)
            seen_nofit += 1
        # (c) less than half of the capacity: no stash
        eng = limited(make, (half // 128 - 2) * 128)
        assert "adjoint=recompute (stash exceeds the memory budget)" in eng.describe(), eng.describe()
        e, f = eng.energy_force(pos)
        assert torch.equal(e, er) and torch.equal(f, fr)
    assert seen_fit >= 2 and seen_nofit >= 1


def test_forward_table_mode_3(weights, monkeypatch):
    """67 ... 126 atoms: the stash forward keeps its Q table in shared memory next to ONE image buffer that the silu'(u) and
    silu'(w) codes of a tile share (table mode 3).  Same arithmetic as the table-less kernel (GNNIS_ONEBUF=0): bit-identical
    results, for whole replicas, split tails and split small batches, in the default and the half-precision arithmetic, and
    with a capacity-limited stash whose edge list does not fit."""
    from gnnimplicitsolvent_b200 import synth
    from gnnimplicitsolvent_b200.engine import Engine
    from oracle import gnn_oracle as O
    nofit = 0
    for N, R, squeeze in ((100, 600, 1.0), (126, 5, 1.0), (67, 300, 1.0), (70, 320, 0.8), (80, 310, 0.7)):
        mol = synth.synth_molecule(N, seed=N)
        pos_h = (synth.conformers(mol["pos"], R, 0.01, seed=N + R) * squeeze).astype(np.float32)
        pos = torch.from_numpy(pos_h).cuda()

        def make():
            eng = Engine(mol["params"], weights, device=0)
            eng.set_replicas(R, [r % 39 for r in range(R)], [synth.SOLVENT_DIELECTRICS[r % 39] for r in range(R)])
            return eng

        eng3 = make()
        assert "tables=3" in eng3.describe(), eng3.describe()
        monkeypatch.setenv("GNNIS_ONEBUF", "0")
        eng0 = make()
        monkeypatch.delenv("GNNIS_ONEBUF")
        assert "tables=0" in eng0.describe(), eng0.describe()
        for kw in ({}, {"fp16": True}, {"delta": True}):
            e3, f3 = eng3.energy_force(pos, **kw)
            e0, f0 = eng0.energy_force(pos, **kw)
            assert torch.equal(e3, e0) and torch.equal(f3, f0), (N, R, kw)
        if squeeze == 1.0 and R <= 8:                      # against the oracle as well (small case)
            sids = [r % 39 for r in range(R)]
            o = O.evaluate(mol["params"], pos_h, sids, [synth.SOLVENT_DIELECTRICS[s_] for s_ in sids],
                           {k: v for k, v in weights.items()})
            e3, f3 = eng3.energy_force(pos)
            np.testing.assert_allclose(e3.cpu().numpy(), o["energy_rep"].numpy(), rtol=E_RTOL, atol=1e-3)
            assert rel_rms(f3.cpu().numpy(), o["forces"].numpy()) < F_RRMS
        # capacity-limited stash that holds half of the capacity but not this edge list: nothing is stashed, the
        # recomputing adjoint runs behind the table-mode-3 forward
        e3, f3 = eng3.energy_force(pos)
        E, cap = eng3.last_edge_count(), R * N * (N - 1)
        hold = -(-(-(-cap // 2)) // 128) * 128
        if hold + 128 <= E:
            monkeypatch.setenv("GNNIS_STASH", "0")
            eng_r = make()
            monkeypatch.delenv("GNNIS_STASH")
            er, fr = eng_r.energy_force(pos)
            monkeypatch.setenv("GNNIS_STASH_GB", repr(((hold + 128) * 768 + 64) / 1e9))
            eng_c = make()
            monkeypatch.delenv("GNNIS_STASH_GB")
            d = eng_c.describe()
            assert f"adjoint=stash up to {hold} edges" in d and "tables=3" in d, d
            for _ in range(2):
                ec, fc = eng_c.energy_force(pos)
                assert torch.equal(ec, er) and torch.equal(fc, fr)
            nofit += 1
    assert nofit >= 1


def test_stash_adjoint_vs_recomputing_adjoint(weights, monkeypatch):
    """Default adjoint (activation stash written by the forward, edge_bwd_st_kernel) against the recomputing one
    (GNNIS_STASH=0, edge_bwd_ws_kernel): identical energies, forces equal to the stash's 16-bit resolution, both inside
    the gates against the reference's goldens; gnnis_describe says which one runs."""
    from gnnimplicitsolvent_b200.engine import Engine
    for name in ("c2_n50_water", "n100_mixed", "mixed_n21"):
        c = load_case(name)
        pos = torch.from_numpy(c["pos"]).cuda()
        eng_s = _engine(c, weights)
        assert "adjoint=stash" in eng_s.describe()
        monkeypatch.setenv("GNNIS_STASH", "0")
        eng_r = _engine(c, weights)
        monkeypatch.delenv("GNNIS_STASH")
        assert "adjoint=recompute" in eng_r.describe()
        es, fs = eng_s.energy_force(pos)
        er, fr = eng_r.energy_force(pos)
        assert torch.equal(es, er)
        assert rel_rms(fs.cpu().numpy(), fr.cpu().numpy()) < 1e-4
        assert rel_rms(fs.cpu().numpy(), c["forces"]) < F_RRMS and rel_rms(fr.cpu().numpy(), c["forces"]) < F_RRMS
        es2, fs2 = eng_s.energy_force(pos)                    # deterministic
        assert torch.equal(fs, fs2)
