"""Runs the PRODUCT sources (kernels + host driver) under the CPU emulation layer of tests/emu and compares with
the oracle.  This exercises indexing, protocol order and the limb algorithms without a GPU; the GPU parity tests
(-m gpu) remain the real gate."""
import numpy as np
import pytest

import oracle_lib

ONE = np.array([8589934590, 6378425256633387010, 11064306276430008309, 1739710354780652911], dtype=np.uint64)


@pytest.fixture(scope="module")
def ctx(pkg, emu_lib):
    c = pkg.ProverContext(8, emu_lib)
    yield c
    c.close()


@pytest.mark.parametrize("logn", [1, 5, 9, 10, 12])
def test_ntt_family(ctx, oracle, logn):
    x = oracle.random_fr(1, 1 << logn)
    for kind in range(4):
        assert np.array_equal(ctx.ntt(kind, x), oracle.ntt(kind, x)), (logn, kind)


@pytest.mark.parametrize("logn", [12, 13, 16, 18])
def test_ntt_register_radix8_kernel(ctx, oracle, monkeypatch, logn):
    """The experimental register-resident radix-8 pass kernel (ZP_NTT_REG=4; passes of 6, 7, 8 and 9 stages): bit-identical
    to the oracle like the default shared-memory kernel."""
    monkeypatch.setenv("ZP_NTT_REG", "4")
    x = oracle.random_fr(1, 1 << logn)
    for kind in range(4):
        assert np.array_equal(ctx.ntt(kind, x), oracle.ntt(kind, x)), (logn, kind)


@pytest.mark.parametrize("logn", [10, 12])
def test_ntt_direct_twiddle_tables(pkg, emu_lib, oracle, monkeypatch, logn):
    """Same transforms when the inter-pass twiddles and the coset-iNTT output factors come from the direct tables."""
    monkeypatch.setenv("ZP_NTT_TW_MIN_LOG", "0")
    c = pkg.ProverContext(8, emu_lib)
    x = oracle.random_fr(1, 1 << logn)
    for kind in range(4):
        assert np.array_equal(c.ntt(kind, x), oracle.ntt(kind, x)), (logn, kind)
    c.close()


def test_scans(ctx, oracle):
    x = oracle.random_fr(3, 3000)
    z = oracle.random_fr(4, 1)[0]
    assert np.array_equal(ctx.poly_eval(x, z), oracle.poly_eval(x, z))
    pp = ctx.prefix_product(x)
    assert np.array_equal(pp[0], ONE)
    assert np.array_equal(pp[1:], oracle.fr_op(2, pp[:-1].copy(), x[:-1].copy()))
    q = ctx.poly_divide(x, z)
    r = oracle.random_fr(5, 1)[0]
    lhs = oracle.fr_op(1, oracle.poly_eval(x, r).reshape(1, 4), oracle.poly_eval(x, z).reshape(1, 4))
    rhs = oracle.fr_op(2, oracle.poly_eval(q, r).reshape(1, 4), oracle.fr_op(1, r.reshape(1, 4), z.reshape(1, 4)))
    assert np.array_equal(lhs, rhs)


@pytest.mark.parametrize("n,wb", [(1, 0), (200, 5), (600, 9), (1024, 0)])
def test_msm(ctx, oracle, n, wb):
    pts, _ = oracle.srs(7, n)
    sc = oracle.random_fr(2, n)
    if n > 100:
        sc[3] = 0
        sc[4] = ONE
        sc[5] = oracle.fr_op(6, ONE.reshape(1, 4))[0]
        pts[10:20] = pts[10]
        sc[10:20] = sc[10]
    assert np.array_equal(ctx.msm_points(pts, sc, wb), oracle.msm(pts, sc))


@pytest.mark.parametrize("n_lookup", [0, 12])
def test_gen_proof_byte_identical_under_emulation(pkg, emu_lib, oracle, n_lookup):
    oc = oracle_lib.OracleCircuit(oracle, 3, 42, 7, n_lookup)  # 3 hashes, N = 2^10
    ref_proof, _ = oc.prove()
    c = pkg.ProverContext(oc.log_n, emu_lib)
    c.load_srs(oc.srs())
    c.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = c.prove(circ).to_words()
    assert np.array_equal(proof, ref_proof)
    assert oc.verify(proof)[0]
    c.close()
    oc.close()


@pytest.mark.parametrize("n_lookup", [0, 12])
def test_gen_proof_with_second_stream_ntts_under_emulation(pkg, emu_lib, oracle, monkeypatch, n_lookup):
    """ZP_NTT_OVERLAP=1 (experiment, off by default): wire and z(X) coset NTTs forked onto the second stream and joined before
    the quotient pass — the fork / join bookkeeping and the skipped jobs give the same proof bytes (the emulation runs the
    streams in program order; the race-freedom of the real streams is what the GPU parity test checks)."""
    monkeypatch.setenv("ZP_NTT_OVERLAP", "1")  # read when the context is created
    oc = oracle_lib.OracleCircuit(oracle, 3, 42, 7, n_lookup)
    ref_proof, _ = oc.prove()
    c = pkg.ProverContext(oc.log_n, emu_lib)
    c.load_srs(oc.srs())
    c.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    assert np.array_equal(c.prove(circ).to_words(), ref_proof)
    assert np.array_equal(c.prove(circ).to_words(), ref_proof)
    c.close()
    oc.close()


def test_gen_proof_with_precomputed_msm_tables(pkg, emu_lib, oracle, monkeypatch):
    """Same proof bytes when the commitments go through the precomputed-window MSM tables (forced on for tiny N)."""
    monkeypatch.setenv("ZP_MSM_PRECOMP_MIN_LOG", "8")
    oc = oracle_lib.OracleCircuit(oracle, 3, 42, 7, 0)
    ref_proof, _ = oc.prove()
    c = pkg.ProverContext(oc.log_n, emu_lib)
    c.load_srs(oc.srs())
    c.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    assert np.array_equal(c.prove(circ).to_words(), ref_proof)
    sc = oracle.random_fr(2, oc.n)
    assert np.array_equal(c.msm(sc), oracle.msm(oc.srs(), sc))
    c.close()
    oc.close()


@pytest.mark.parametrize("rounds", [1, 3])
def test_msm_batch_affine_rounds(pkg, emu_lib, oracle, monkeypatch, rounds):
    """Batch-affine pre-reduction (pairwise affine additions with a shared batch inversion) gives the same point; with
    repeated input points (x1 == x2 pairs) it must detect the degenerate pair and fall back to the exact XYZZ path."""
    monkeypatch.setenv("ZP_MSM_BA_ROUNDS", str(rounds))
    monkeypatch.setenv("ZP_MSM_BA_MIN_LOG", "4")
    ctx = pkg.ProverContext(8, emu_lib)
    n = 700
    pts, _ = oracle.srs(7, n)
    sc = oracle.random_fr(2, n)
    sc[3] = 0
    assert np.array_equal(ctx.msm_points(pts, sc, 6), oracle.msm(pts, sc))
    pts2, sc2 = pts.copy(), sc.copy()
    pts2[10:40] = pts2[10]
    sc2[10:40] = sc2[10]
    assert np.array_equal(ctx.msm_points(pts2, sc2, 6), oracle.msm(pts2, sc2))
    # skewed scalars over distinct points: 500 of 700 scalars are 3, so one bucket run has 250 pairs (written by the whole
    # warp in ba_slots_kernel) and is still split into more than 8 work segments afterwards (msm_fold_kernel's work list)
    sc3 = sc.copy()
    sc3[100:600] = _fr_small(oracle, [3])[0]
    assert np.array_equal(ctx.msm_points(pts, sc3, 6), oracle.msm(pts, sc3))
    ctx.close()


def test_msm_batch_members(pkg, emu_lib, oracle, monkeypatch):
    """Several scalar vectors in one MSM pipeline (with and without the precomputed table / batch-affine rounds)."""
    n = 600
    srs, _ = oracle.srs(7, 1024)
    pts = srs[:n].copy()
    sc = np.stack([oracle.random_fr(40 + j, n) for j in range(3)])
    sc[1] = 0
    for env in ({}, {"ZP_MSM_PRECOMP_MIN_LOG": "8", "ZP_MSM_BA_ROUNDS": "2", "ZP_MSM_BA_MIN_LOG": "4"}):
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        c = pkg.ProverContext(10, emu_lib)
        c.load_srs(srs)
        out = c.msm_batch(sc)
        for j in range(3):
            assert np.array_equal(out[j], oracle.msm(pts, sc[j].copy())), (env, j)
        c.close()


def test_gen_proof_with_batch_affine_and_tables(pkg, emu_lib, oracle, monkeypatch):
    monkeypatch.setenv("ZP_MSM_PRECOMP_MIN_LOG", "8")
    monkeypatch.setenv("ZP_MSM_BA_ROUNDS", "2")
    monkeypatch.setenv("ZP_MSM_BA_MIN_LOG", "4")
    oc = oracle_lib.OracleCircuit(oracle, 3, 42, 7, 0)
    ref_proof, _ = oc.prove()
    c = pkg.ProverContext(oc.log_n, emu_lib)
    c.load_srs(oc.srs())
    c.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    assert np.array_equal(c.prove(circ).to_words(), ref_proof)
    c.close()
    oc.close()


def test_edge_cases(ctx, oracle):
    """Empty / single-element / ragged inputs of the operator entry points."""
    pts, _ = oracle.srs(7, 4)
    sc = oracle.random_fr(2, 4)
    inf = ctx.msm_points(pts[:0].copy(), sc[:0].copy())  # empty MSM = identity in the FFI encoding (0, Mont(1))
    assert not inf[:6].any() and inf[6] == 0x760900000002fffd
    assert np.array_equal(ctx.msm_points(pts[:1].copy(), sc[:1].copy()), oracle.msm(pts[:1].copy(), sc[:1].copy()))
    for n in [1, 2]:
        x = oracle.random_fr(1, n)
        for kind in range(4):
            assert np.array_equal(ctx.ntt(kind, x), oracle.ntt(kind, x))
    z = oracle.random_fr(4, 1)[0]
    for n in [1, 5, 33, 64, 65]:
        x = oracle.random_fr(3, n)
        assert np.array_equal(ctx.poly_eval(x, z), oracle.poly_eval(x, z))
        assert np.array_equal(ctx.prefix_product(x)[0], ONE)


def _fr_small(oracle, vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    a[:, 0] = vals
    return oracle.fr_op(5, a)


def test_combine_split_device(ctx, oracle, pkg):
    """plookup combine_split on the device vs the oracle (multiset.rs:131-176), incl. the doc example and the error."""
    h1, h2 = ctx.combine_split(_fr_small(oracle, [2, 4, 1, 3]), _fr_small(oracle, [2, 3, 3, 2]))
    assert np.array_equal(h1, _fr_small(oracle, [2, 2, 1, 3])) and np.array_equal(h2, _fr_small(oracle, [2, 4, 3, 3]))
    rng = np.random.default_rng(5)
    n = 1000
    tv = rng.integers(0, 300, n)          # many repeated table values, arbitrary order
    fv = rng.choice(tv, n)
    t, f = _fr_small(oracle, tv), _fr_small(oracle, fv)
    ok, o1, o2 = oracle.combine_split(t, f)
    assert ok
    h1, h2 = ctx.combine_split(t, f)
    assert np.array_equal(h1, o1) and np.array_equal(h2, o2)
    with pytest.raises(pkg.ZprizeError, match="ElementNotIndexed"):
        ctx.combine_split(_fr_small(oracle, [2, 4, 1, 3]), _fr_small(oracle, [2, 3, 5, 2]))
