"""Checks the oracle against the reference's OWN sources compiled into oracle/_ref (blst, strobe.cpp).
Skipped when oracle/_ref is absent (it is built here from /root/reference and travels with the snapshot)."""
import ctypes
import random

import numpy as np
import pytest

import oracle_lib
from oracle_lib import _p

u64p = oracle_lib.u64p


@pytest.fixture(scope="module")
def blst():
    lib = oracle_lib.ref_lib("libref_blst.so")
    if lib is None:
        pytest.skip("oracle/_ref/libref_blst.so not built")
    return lib


@pytest.fixture(scope="module")
def strobe():
    lib = oracle_lib.ref_lib("libref_strobe.so")
    if lib is None:
        pytest.skip("oracle/_ref/libref_strobe.so not built")
    return lib


def test_fr_arithmetic_vs_blst(oracle, blst):
    n = 2000
    a, b = oracle.random_fr(21, n), oracle.random_fr(22, n)
    mul, add, sub = oracle.fr_op(2, a, b), oracle.fr_op(0, a, b), oracle.fr_op(1, a, b)
    inv = oracle.fr_op(3, a[:50].copy())
    canon = oracle.fr_op(4, a)
    r = np.zeros(4, dtype=np.uint64)
    for i in range(n):
        blst.blst_fr_mul(_p(r), _p(a[i].copy()), _p(b[i].copy()))
        assert np.array_equal(r, mul[i])
        blst.blst_fr_add(_p(r), _p(a[i].copy()), _p(b[i].copy()))
        assert np.array_equal(r, add[i])
        blst.blst_fr_sub(_p(r), _p(a[i].copy()), _p(b[i].copy()))
        assert np.array_equal(r, sub[i])
        blst.blst_fr_from(_p(r), _p(a[i].copy()))  # Montgomery -> canonical
        assert np.array_equal(r, canon[i])
    for i in range(50):
        blst.blst_fr_eucl_inverse(_p(r), _p(a[i].copy()))
        assert np.array_equal(r, inv[i])


def _random_fq(oracle, seed, n):
    rnd = random.Random(seed)
    q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab", 16)
    vals = np.zeros((n, 6), dtype=np.uint64)
    for i in range(n):
        v = rnd.randrange(q)
        for j in range(6):
            vals[i, j] = (v >> (64 * j)) & 0xffffffffffffffff
    return oracle.fq_op(5, vals)  # canonical -> Montgomery


def test_fq_arithmetic_vs_blst(oracle, blst):
    n = 1000
    a, b = _random_fq(oracle, 1, n), _random_fq(oracle, 2, n)
    mul, add, sub = oracle.fq_op(2, a, b), oracle.fq_op(0, a, b), oracle.fq_op(1, a, b)
    r = np.zeros(6, dtype=np.uint64)
    for i in range(n):
        blst.blst_fp_mul(_p(r), _p(a[i].copy()), _p(b[i].copy()))
        assert np.array_equal(r, mul[i])
        blst.blst_fp_add(_p(r), _p(a[i].copy()), _p(b[i].copy()))
        assert np.array_equal(r, add[i])
        blst.blst_fp_sub(_p(r), _p(a[i].copy()), _p(b[i].copy()))
        assert np.array_equal(r, sub[i])


def test_g1_generator_and_scalar_mul_vs_blst(oracle, blst):
    blst.blst_p1_affine_generator.restype = u64p
    g_ref = np.ctypeslib.as_array(blst.blst_p1_affine_generator(), shape=(12,)).copy()
    g = np.zeros(12, dtype=np.uint64)
    oracle.lib.zpo_g1_generator(_p(g))
    assert np.array_equal(g, g_ref)
    blst.blst_p1_generator.restype = u64p
    gj = np.ctypeslib.as_array(blst.blst_p1_generator(), shape=(18,)).copy()
    sc = oracle.random_fr(5, 8)
    for i in range(8):
        canon = oracle.fr_op(4, sc[i:i + 1])[0].copy()
        out = np.zeros(18, dtype=np.uint64)
        blst.blst_p1_mult(_p(out), _p(gj), ctypes.cast(_p(canon), ctypes.c_void_p), ctypes.c_size_t(255))
        aff = np.zeros(12, dtype=np.uint64)
        blst.blst_p1_to_affine(_p(aff), _p(out))
        assert np.array_equal(aff, oracle.g1_mul(g, sc[i]))


def test_msm_vs_blst_pippenger(oracle, blst):
    n = 300
    pts, _ = oracle.srs(7, n)
    sc = oracle.random_fr(2, n)
    canon = oracle.fr_op(4, sc)
    blst.blst_p1s_mult_pippenger_scratch_sizeof.restype = ctypes.c_size_t
    blst.blst_p1s_mult_pippenger_scratch_sizeof.argtypes = [ctypes.c_size_t]
    scratch = ctypes.create_string_buffer(blst.blst_p1s_mult_pippenger_scratch_sizeof(n))
    pp = (ctypes.c_void_p * 2)(pts.ctypes.data, None)
    sp = (ctypes.c_void_p * 2)(canon.ctypes.data, None)
    out = np.zeros(18, dtype=np.uint64)
    blst.blst_p1s_mult_pippenger(_p(out), pp, ctypes.c_size_t(n), sp, ctypes.c_size_t(255), scratch)
    aff = np.zeros(12, dtype=np.uint64)
    blst.blst_p1_to_affine(_p(aff), _p(out))
    assert np.array_equal(aff, oracle.msm(pts, sc))


def test_transcript_vs_reference_strobe(oracle, strobe):
    """Replays a prover-shaped transcript on the reference's strobe.cpp and on the oracle."""
    rnd = random.Random(3)
    ops = [("append", b"pi", bytes(rnd.randrange(256) for _ in range(48)))]
    for lab in [b"w_l", b"w_r", b"w_o", b"w_4"]:
        ops.append(("append", lab, bytes(rnd.randrange(256) for _ in range(48))))
    ops += [("challenge", b"zeta", 31), ("append", b"zeta", bytes(32)), ("append", b"f", bytes(48))]
    for lab in [b"beta", b"gamma", b"delta", b"epsilon"]:
        ops += [("challenge", lab, 31), ("append", lab, bytes(rnd.randrange(256) for _ in range(32)))]
    ops += [("challenge", b"range separation challenge", 31), ("append", b"range seperation challenge", bytes(32))]
    ops += [("append", b"big", bytes(rnd.randrange(256) for _ in range(700))), ("challenge", b"aggregate_witness", 31),
            ("challenge", b"aggregate_witness", 31), ("challenge", b"wide", 400)]
    script, total = oracle_lib.encode_script(ops)
    ref_out = ctypes.create_string_buffer(total)
    strobe.ref_transcript_script.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    strobe.ref_transcript_script(b"Merkle tree", script, len(script), ref_out)
    assert oracle.transcript_script(b"Merkle tree", ops) == ref_out.raw
