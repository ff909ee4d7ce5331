"""The oracle's prover restatement against its verifier restatement, plus primitives' defining properties."""
import numpy as np
import pytest

import oracle_lib

ONE = np.array([8589934590, 6378425256633387010, 11064306276430008309, 1739710354780652911], dtype=np.uint64)


@pytest.fixture(scope="module")
def circuit4(oracle):
    oc = oracle_lib.OracleCircuit(oracle, 4, 42, 7, 0)
    yield oc
    oc.close()


def test_height4_shape(circuit4):
    # 1 zero gate + 3 blinding rows + 7 hashes * 193 gates + 1 public-input gate
    assert circuit4.cs_n == 4 + 7 * 193 + 1 == 1356
    assert circuit4.log_n == 11
    assert circuit4.satisfied()


def test_prove_then_verify_and_tamper(circuit4):
    proof, _ = circuit4.prove()
    ok, detail = circuit4.verify(proof)
    assert ok and detail == 3
    proof2, _ = circuit4.prove()
    assert np.array_equal(proof, proof2)  # deterministic: no prover-side randomness (prover.rs:293-317)
    for word in [0, 12 * 4 + 3, 12 * 17 + 1, 12 * 19 + 2, 12 * 19 + 4 * 7]:
        bad = proof.copy()
        bad[word] ^= 1
        assert not circuit4.verify(bad)[0]
    # f, h1, h2, t7, t8 are the identity for the Merkle circuit (merkle-tree/src/main.rs:112-123)
    fq_one = np.array([0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745,
                       0x5c071a97a256ec6d, 0x15f65ec3fa80e493], dtype=np.uint64)
    for idx in [5, 6, 7, 15, 16]:
        c = proof[12 * idx:12 * idx + 12]
        assert not c[:6].any() and np.array_equal(c[6:], fq_one)


def test_lookup_variant_verifies(oracle):
    oc = oracle_lib.OracleCircuit(oracle, 4, 43, 7, 24)
    assert oc.lookup_len == 16 and oc.satisfied()
    proof, _ = oc.prove()
    assert oc.verify(proof)[0]
    # with real lookups f / h1 / h2 are NOT the identity
    assert proof[12 * 5:12 * 5 + 6].any()
    oc.close()


def test_combine_split_paper_example(oracle):
    # multiset.rs:118-130 doc example: t = {2,4,1,3}, f = {2,3,3,2} -> h1 = {2,2,1,3}, h2 = {2,4,3,3}
    def fr(vals):
        a = np.zeros((len(vals), 4), dtype=np.uint64)
        a[:, 0] = vals
        return oracle.fr_op(5, a)
    ok, h1, h2 = oracle.combine_split(fr([2, 4, 1, 3]), fr([2, 3, 3, 2]))
    assert ok
    assert np.array_equal(h1, fr([2, 2, 1, 3])) and np.array_equal(h2, fr([2, 4, 3, 3]))
    ok, _, _ = oracle.combine_split(fr([2, 4, 1, 3]), fr([2, 3, 5, 2]))
    assert not ok  # Error::ElementNotIndexed


@pytest.mark.parametrize("logn", [1, 4, 9])
def test_ntt_defining_property(oracle, logn):
    n = 1 << logn
    x = oracle.random_fr(9, n)
    w, wi, ni = (np.zeros(4, np.uint64) for _ in range(3))
    oracle.lib.zpo_fr_root_of_unity(logn, oracle_lib._p(w), oracle_lib._p(wi), oracle_lib._p(ni))
    fx = oracle.ntt(0, x)
    pt = ONE.copy()
    for i in range(min(n, 8)):
        assert np.array_equal(fx[i], oracle.poly_eval(x, pt))  # evals[i] = p(omega^i)
        pt = oracle.fr_op(2, pt.reshape(1, 4), w.reshape(1, 4))[0]
    assert np.array_equal(oracle.ntt(1, fx), x)
    cx = oracle.ntt(2, x)
    g = np.array([64424509425, 1721329240476523535, 18418692815241631664, 3824455624000121028], dtype=np.uint64)
    assert np.array_equal(cx[0], oracle.poly_eval(x, g))  # coset_fft evaluates at g * omega^i, g = 7
    assert np.array_equal(oracle.ntt(3, cx), x)


def test_known_tau_commitment(circuit4, oracle):
    co = circuit4.pk_coeffs()[1]
    srs = circuit4.srs()
    assert np.array_equal(oracle.msm(srs, co), circuit4.commit_with_tau(co))  # commit(p) == [p(tau)] G
