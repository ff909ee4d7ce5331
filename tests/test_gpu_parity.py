"""GPU parity tests proper: every call goes through the C-ABI of libzprize_b200.so on a real device and is
compared bit for bit with the CPU oracle (integer arithmetic: the bar is exact equality)."""
import ctypes

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu

ONE = np.array([8589934590, 6378425256633387010, 11064306276430008309, 1739710354780652911], dtype=np.uint64)


@pytest.fixture(scope="module")
def ctx16(pkg, gpu_lib):
    c = pkg.ProverContext(16, gpu_lib)
    yield c
    c.close()


@pytest.mark.parametrize("logn", [6, 9, 10, 13, 16, 18, 19])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_ntt_matches_oracle(ctx16, oracle, logn, kind):
    x = oracle.random_fr(1, 1 << logn)
    assert np.array_equal(ctx16.ntt(kind, x), oracle.ntt(kind, x))


def test_ntt_roundtrip_and_linearity_large(ctx16, oracle):
    # size-independent properties at 2^22 (the HEIGHT=15 domain): iNTT(NTT(x)) = x, NTT(x + y) = NTT(x) + NTT(y)
    n = 1 << 22
    x, y = oracle.random_fr(11, n), oracle.random_fr(12, n)
    fx = ctx16.ntt(0, x)
    assert np.array_equal(ctx16.ntt(1, fx), x)
    fy = ctx16.ntt(0, y)
    assert np.array_equal(ctx16.ntt(0, oracle.fr_op(0, x, y)), oracle.fr_op(0, fx, fy))
    cx = ctx16.ntt(2, x)
    assert np.array_equal(ctx16.ntt(3, cx), x)
    # spot-check 4 evaluations against Horner on the CPU
    w, _, _ = np.zeros(4, np.uint64), None, None
    fwd, inv, ninv = np.zeros(4, np.uint64), np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    oracle.lib.zpo_fr_root_of_unity(22, oracle_lib._p(fwd), oracle_lib._p(inv), oracle_lib._p(ninv))
    pt = ONE.copy()
    for i in range(3):
        assert np.array_equal(fx[i], oracle.poly_eval(x, pt))
        pt = oracle.fr_op(2, pt.reshape(1, 4), fwd.reshape(1, 4))[0]


def test_scan_eval_divide(ctx16, oracle):
    n = 70001
    x = oracle.random_fr(3, n)
    z = oracle.random_fr(4, 1)[0]
    assert np.array_equal(ctx16.poly_eval(x, z), oracle.poly_eval(x, z))
    q = ctx16.poly_divide(x, z)
    # p(X) - p(z) = q(X) (X - z): check at a random point
    r = oracle.random_fr(5, 1)[0]
    lhs = oracle.fr_op(1, oracle.poly_eval(x, r).reshape(1, 4), oracle.poly_eval(x, z).reshape(1, 4))
    rhs = oracle.fr_op(2, oracle.poly_eval(q, r).reshape(1, 4), oracle.fr_op(1, r.reshape(1, 4), z.reshape(1, 4)))
    assert np.array_equal(lhs, rhs)
    pp = ctx16.prefix_product(x)
    assert np.array_equal(pp[0], ONE)
    # telescoping: pp[i+1] = pp[i] * x[i]
    assert np.array_equal(pp[1:], oracle.fr_op(2, pp[:-1].copy(), x[:-1].copy()))


@pytest.mark.parametrize("n,wb", [(1, 0), (37, 0), (1 << 10, 0), (1 << 10, 5), (1 << 12, 0), (1 << 14, 13), (1 << 16, 0)])
def test_msm_matches_oracle(ctx16, oracle, n, wb):
    pts, _ = oracle.srs(7, n)
    sc = oracle.random_fr(2, n)
    assert np.array_equal(ctx16.msm_points(pts, sc, wb), oracle.msm(pts, sc))


def test_msm_adversarial_scalars(ctx16, oracle):
    n = 1 << 10
    pts, _ = oracle.srs(7, n)
    sc = oracle.random_fr(2, n)
    minus_one = oracle.fr_op(6, ONE.reshape(1, 4))[0]
    sc[0:64] = 0            # zeros
    sc[64:128] = ONE        # ones
    sc[128:192] = minus_one  # r - 1: every window digit at its extreme
    pts[200:264] = pts[200]  # repeated points (forces the doubling branch of the bucket add)
    sc[200:264] = sc[200]
    assert np.array_equal(ctx16.msm_points(pts, sc), oracle.msm(pts, sc))
    z = np.zeros_like(sc)
    out = ctx16.msm_points(pts, z)
    assert not out[:6].any() and np.array_equal(out[6:], oracle_fq_one(oracle))


def oracle_fq_one(oracle):
    m, one, rr = (np.zeros(6, np.uint64) for _ in range(3))
    inv = ctypes.c_uint64()
    oracle.lib.zpo_fq_constants(oracle_lib._p(m), oracle_lib._p(one), oracle_lib._p(rr),
                                ctypes.cast(ctypes.byref(inv), oracle_lib.u64p))
    return one


def test_device_srs_matches_oracle(pkg, gpu_lib, oracle):
    ctx = pkg.ProverContext(10, gpu_lib)
    pts, tau = oracle.srs(7, 1 << 10)
    ctx.generate_srs(tau)
    assert np.array_equal(ctx.read_srs(), pts)
    ctx.close()


def _prove_and_compare(pkg, lib, oracle, height, n_lookup, use_host_pk):
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, n_lookup)
    ref_proof, _ = oc.prove()
    ctx = pkg.ProverContext(oc.log_n, lib)
    ctx.load_srs(oc.srs())
    keep = None
    if use_host_pk:
        co, ev = oc.pk_coeffs(), oc.pk_evals()
        names = pkg.PK_POLY_NAMES + pkg.PK_SIGMA_NAMES
        keep = (co, ev, oc.tables(), oc.linear_evaluations(), oc.v_h_coset_8n())
        pk = pkg.make_prover_key(dict(zip(names, co)), dict(zip(names, ev)), keep[2], keep[3], keep[4])
        ctx.load_pk(pk)
    else:
        ctx.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = ctx.prove(circ).to_words()
    assert np.array_equal(proof, ref_proof)
    ok, _ = oc.verify(proof)
    assert ok
    # verifier key computed on the device equals the oracle's
    ctx.close()
    oc.close()


@pytest.mark.parametrize("height,n_lookup,use_host_pk", [(4, 0, False), (4, 0, True), (4, 24, False), (4, 24, True),
                                                        (6, 0, False)])
def test_gen_proof_byte_identical(pkg, gpu_lib, oracle, height, n_lookup, use_host_pk):
    _prove_and_compare(pkg, gpu_lib, oracle, height, n_lookup, use_host_pk)


def test_drop_in_gen_proof_symbol(pkg, gpu_lib, oracle):
    """The reference's own FFI call: by-value structs in, ProofC by value out (lib.rs:237-239)."""
    oc = oracle_lib.OracleCircuit(oracle, 4, 42, 7, 0)
    ref_proof, _ = oc.prove()
    names = pkg.PK_POLY_NAMES + pkg.PK_SIGMA_NAMES
    co, ev, tb = oc.pk_coeffs(), oc.pk_evals(), oc.tables()
    le, vh = oc.linear_evaluations(), oc.v_h_coset_8n()
    pk = pkg.make_prover_key(dict(zip(names, co)), dict(zip(names, ev)), tb, le, vh)
    # the reference passes dangling pointers for the coefficient arrays of identically-zero selectors
    for nm in ["q_m", "range_selector", "logic_selector", "fixed_group_add_selector", "variable_group_add_selector", "q_lookup"]:
        setattr(pk, nm + "_coeffs", ctypes.cast(0xdead0000, pkg.u64p))
    srs = oc.srs()
    ck = pkg.CommitKeyC()
    ck.powers_of_g = pkg.as_u64p(srs)
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    for _ in range(2):  # second call hits the resident-key cache
        proof = pkg.gen_proof(circ, pk, ck, gpu_lib)
        assert np.array_equal(proof.to_words(), ref_proof)
    oc.close()


def _ffi_inputs(pkg, oc):
    names = pkg.PK_POLY_NAMES + pkg.PK_SIGMA_NAMES
    keep = [oc.pk_coeffs(), oc.pk_evals(), oc.tables(), oc.linear_evaluations(), oc.v_h_coset_8n(), oc.srs()]
    pk = pkg.make_prover_key(dict(zip(names, keep[0])), dict(zip(names, keep[1])), keep[2], keep[3], keep[4])
    ck = pkg.CommitKeyC()
    ck.powers_of_g = pkg.as_u64p(keep[5])
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    return circ, pk, ck, keep


def test_gen_proof_cache_is_keyed_on_content(pkg, gpu_lib, oracle):
    """The reference's harness clones the prover key for every proof and rebuilds powers_of_g per call
    (benches/pnp_bench.rs:62-118, prover.rs:700-711): same key at new addresses must NOT re-upload, and a different key
    arriving in recycled buffers must not be served from the cache."""
    import time
    oc_a = oracle_lib.OracleCircuit(oracle, 6, 42, 7, 0)
    oc_b = oracle_lib.OracleCircuit(oracle, 6, 42, 7, 8)  # 8 plookup rows + a table: other selectors, sigmas and tables, same N
    assert oc_b.log_n == oc_a.log_n and not np.array_equal(oc_a.selector_evals()[15], oc_b.selector_evals()[15])
    ref_a, _ = oc_a.prove()
    ref_b, _ = oc_b.prove()
    gpu_lib.zp_gen_proof_invalidate()
    circ, pk, ck, keep = _ffi_inputs(pkg, oc_a)
    t0 = time.perf_counter()
    assert np.array_equal(pkg.gen_proof(circ, pk, ck, gpu_lib).to_words(), ref_a)
    cold = time.perf_counter() - t0
    # "pk.clone()": fresh buffers, same content
    circ2, pk2, ck2, keep2 = _ffi_inputs(pkg, oc_a)
    assert keep2[1][1].ctypes.data != keep[1][1].ctypes.data
    t0 = time.perf_counter()
    assert np.array_equal(pkg.gen_proof(circ2, pk2, ck2, gpu_lib).to_words(), ref_a)
    warm = time.perf_counter() - t0
    assert warm < cold, (warm, cold)
    # another key written INTO the buffers of the first one (same addresses)
    circ_b, pk_b, ck_b, keep_b = _ffi_inputs(pkg, oc_b)
    for dst, src in zip(keep[0] + keep[1] + keep[2], keep_b[0] + keep_b[1] + keep_b[2]):
        dst[...] = src
    proof = pkg.gen_proof(circ_b, pk, ck, gpu_lib).to_words()
    assert np.array_equal(proof, ref_b)
    gpu_lib.zp_gen_proof_invalidate()
    oc_a.close()
    oc_b.close()


def test_verifier_key_and_known_tau(pkg, gpu_lib, oracle):
    oc = oracle_lib.OracleCircuit(oracle, 4, 42, 7, 0)
    ctx = pkg.ProverContext(oc.log_n, gpu_lib)
    ctx.load_srs(oc.srs())
    ctx.preprocess(oc.selector_evals(), oc.tables())
    vk = ctx.verifier_key()
    co = oc.pk_coeffs()
    for i in [1, 3, 5, 6, 9, 15, 18]:
        assert np.array_equal(vk[i], oc.commit_with_tau(co[i]))  # commit(p) == [p(tau)] G
    ctx.close()
    oc.close()


def test_precomputed_table_msm_matches_oracle(pkg, gpu_lib, oracle):
    """MSM over the resident SRS through the precomputed window tables (n >= 2^16) vs the CPU oracle."""
    ctx = pkg.ProverContext(16, gpu_lib)
    pts, tau = oracle.srs(7, 1 << 16)
    ctx.load_srs(pts)
    sc = oracle.random_fr(2, 1 << 16)
    sc[:100] = 0
    assert np.array_equal(ctx.msm(sc), oracle.msm(pts, sc))
    ctx.close()


def test_edge_cases_on_device(ctx16, oracle):
    pts, _ = oracle.srs(7, 4)
    sc = oracle.random_fr(2, 4)
    inf = ctx16.msm_points(pts[:0].copy(), sc[:0].copy())
    assert not inf[:6].any() and np.array_equal(inf[6:], oracle_fq_one(oracle))
    for n in [1, 2]:
        x = oracle.random_fr(1, n)
        for kind in range(4):
            assert np.array_equal(ctx16.ntt(kind, x), oracle.ntt(kind, x))
    z = oracle.random_fr(4, 1)[0]
    for n in [1, 5, 33, 64, 65]:
        x = oracle.random_fr(3, n)
        assert np.array_equal(ctx16.poly_eval(x, z), oracle.poly_eval(x, z))


def test_largest_supported_ntt_roundtrip(ctx16, oracle):
    """2^25 = the 8N extended domain of HEIGHT=15: coset NTT then coset iNTT is the identity; plain NTT is linear."""
    n = 1 << 25
    x = oracle.random_fr(21, n)
    assert np.array_equal(ctx16.ntt(3, ctx16.ntt(2, x)), x)


def test_batch_affine_msm_on_device(pkg, gpu_lib, oracle, monkeypatch):
    monkeypatch.setenv("ZP_MSM_BA_ROUNDS", "3")
    monkeypatch.setenv("ZP_MSM_BA_MIN_LOG", "10")
    ctx = pkg.ProverContext(10, gpu_lib)
    n = 1 << 14
    pts, _ = oracle.srs(7, n)
    sc = oracle.random_fr(2, n)
    assert np.array_equal(ctx.msm_points(pts, sc), oracle.msm(pts, sc))
    pts[100:164] = pts[100]  # equal points in one bucket: degenerate pair -> exact fallback path
    sc[100:164] = sc[100]
    assert np.array_equal(ctx.msm_points(pts, sc), oracle.msm(pts, sc))
    ctx.close()


@pytest.mark.parametrize("logn,k", [(10, 3), (16, 5), (18, 8)])
def test_msm_batch_matches_oracle(pkg, gpu_lib, oracle, logn, k):
    """k scalar vectors over the same SRS points through ONE MSM pipeline (the path of the prover's independent
    commitments): every member equals the oracle's single MSM; members include an all-zero vector and repeated scalars.
    2^16 and 2^18 go through the precomputed window table and the batch-affine rounds."""
    n = 1 << logn
    ctx = pkg.ProverContext(logn, gpu_lib)
    pts, tau = oracle.srs(7, n)
    ctx.generate_srs(tau)
    sc = np.stack([oracle.random_fr(30 + j, n) for j in range(k)])
    sc[1] = 0
    sc[2, : n // 2] = sc[2, 0]
    out = ctx.msm_batch(sc)
    for j in range(k):
        if logn <= 16 or j < 3:
            assert np.array_equal(out[j], oracle.msm(pts, sc[j].copy())), (logn, j)
    # same members one by one through the single-MSM entry point
    for j in (0, k - 1):
        assert np.array_equal(out[j], ctx.msm(sc[j].copy()))
    ctx.close()


def test_msm_full_size_trapdoor_identity(pkg, gpu_lib, oracle):
    """BASELINE full sizes through a size-independent property: with the known-trapdoor SRS [tau^i] G,
    MSM(s) = [sum_i s_i tau^i] G.  2^22 is the HEIGHT=15 commitment (precomputed table, 4 batch-affine rounds, batch of 2);
    2^24 runs in an operator-only context (no 8N domain)."""
    for logn, k in [(22, 2), (24, 1)]:
        n = 1 << logn
        ctx = pkg.ProverContext(logn, gpu_lib)
        tau = oracle.random_fr(7, 1)[0]
        ctx.generate_srs(tau)
        g = ctx.read_srs(1)[0]
        sc = np.stack([oracle.random_fr(50 + j, n) for j in range(k)])
        out = ctx.msm_batch(sc)
        for j in range(k):
            assert np.array_equal(out[j], oracle.g1_mul(g, oracle.poly_eval(sc[j], tau))), (logn, j)
        ctx.close()


def test_combine_split_on_device(ctx16, oracle, pkg):
    def small(vals):
        a = np.zeros((len(vals), 4), dtype=np.uint64)
        a[:, 0] = vals
        return oracle.fr_op(5, a)
    rng = np.random.default_rng(6)
    n = 1 << 14
    tv = rng.integers(0, 5000, n)
    fv = rng.choice(tv, n)
    t, f = small(tv), small(fv)
    ok, o1, o2 = oracle.combine_split(t, f)
    assert ok
    h1, h2 = ctx16.combine_split(t, f)
    assert np.array_equal(h1, o1) and np.array_equal(h2, o2)
    with pytest.raises(pkg.ZprizeError, match="ElementNotIndexed"):
        ctx16.combine_split(small([2, 4, 1, 3]), small([2, 3, 5, 2]))
