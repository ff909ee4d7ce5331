"""The known-answer vectors the reference's OWN unit tests hold for operators on the hot path, replayed against the
oracle (CPU) and the product (emulated here, on the device under -m gpu):

  * MultiSet::combine_split      "Prize 1B/plonk-core/src/lookup/multiset.rs":335-393  (t of 7, f of 11 elements)
                                 and the doc example :118-130 (t {2,4,1,3}, f {2,3,3,2})
  * sigma permutations/encodings "…/permutation/mod.rs":970-1190 (two wire maps with explicit WireData + K_i w^j)
  * lc (MultiSet::compress)      "…/util.rs":303-321 (a + b c + c c^2 + d c^3 + e c^4)
"""
import ctypes

import numpy as np
import pytest

import oracle_lib
from oracle_lib import _p


def fr_small(oracle, vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    a[:, 0] = vals
    return oracle.fr_op(5, a)


# lookup/multiset.rs:335-393
T7 = [0, 1, 2, 3, 4, 5, 6]
F11 = [3, 6, 0, 5, 4, 3, 2, 0, 0, 1, 2]
EVENS = [0, 0, 1, 2, 2, 3, 4, 5, 6]
ODDS = [0, 0, 1, 2, 3, 3, 4, 5, 6]


def _oracle_cs(oracle, t, f):
    nt, nf = t.shape[0], f.shape[0]
    h1 = np.zeros(((nt + nf + 1) // 2, 4), dtype=np.uint64)
    h2 = np.zeros(((nt + nf) // 2, 4), dtype=np.uint64)
    oracle.lib.zpo_multiset_combine_split.argtypes = [ctypes.c_size_t, oracle_lib.u64p, ctypes.c_size_t, oracle_lib.u64p,
                                                      oracle_lib.u64p, oracle_lib.u64p]
    ok = oracle.lib.zpo_multiset_combine_split(nt, _p(t), nf, _p(f), _p(h1), _p(h2))
    return ok, h1, h2


def test_oracle_combine_split_reference_vector(oracle):
    ok, h1, h2 = _oracle_cs(oracle, fr_small(oracle, T7), fr_small(oracle, F11))
    assert ok == 1
    assert np.array_equal(h1, fr_small(oracle, EVENS)) and np.array_equal(h2, fr_small(oracle, ODDS))
    ok, h1, h2 = _oracle_cs(oracle, fr_small(oracle, [2, 4, 1, 3]), fr_small(oracle, [2, 3, 3, 2]))
    assert np.array_equal(h1, fr_small(oracle, [2, 2, 1, 3])) and np.array_equal(h2, fr_small(oracle, [2, 4, 3, 3]))
    assert _oracle_cs(oracle, fr_small(oracle, T7), fr_small(oracle, [8]))[0] == 0  # Error::ElementNotIndexed


def _product_combine_split(pkg, lib, oracle):
    ctx = pkg.ProverContext(8, lib)
    h1, h2 = ctx.multiset_combine_split(fr_small(oracle, T7), fr_small(oracle, F11))
    assert np.array_equal(h1, fr_small(oracle, EVENS)) and np.array_equal(h2, fr_small(oracle, ODDS))
    h1, h2 = ctx.multiset_combine_split(fr_small(oracle, [2, 4, 1, 3]), fr_small(oracle, [2, 3, 3, 2]))
    assert np.array_equal(h1, fr_small(oracle, [2, 2, 1, 3])) and np.array_equal(h2, fr_small(oracle, [2, 4, 3, 3]))
    with pytest.raises(pkg.ZprizeError, match="ElementNotIndexed"):
        ctx.multiset_combine_split(fr_small(oracle, T7), fr_small(oracle, [8]))
    # odd total: h1 gets the extra element
    rng = np.random.default_rng(3)
    tv = rng.integers(0, 50, 101)
    fv = rng.choice(tv, 300)
    t, f = fr_small(oracle, tv), fr_small(oracle, fv)
    ok, o1, o2 = _oracle_cs(oracle, t, f)
    h1, h2 = ctx.multiset_combine_split(t, f)
    assert ok == 1 and h1.shape[0] == 201 and h2.shape[0] == 200
    assert np.array_equal(h1, o1) and np.array_equal(h2, o2)
    ctx.close()


def test_emulated_combine_split_reference_vector(pkg, emu_lib, oracle):
    _product_combine_split(pkg, emu_lib, oracle)


@pytest.mark.gpu
def test_gpu_combine_split_reference_vector(pkg, gpu_lib, oracle):
    _product_combine_split(pkg, gpu_lib, oracle)


# ---- permutation/mod.rs:970-1190.  WireData codes: (wire << 28) | gate, wire 0 = Left, 1 = Right, 2 = Output, 3 = Fourth
L, R, O, F = 0, 1, 2, 3


def wd(wire, gate):
    return (wire << 28) | gate


SIGMA_CASES = [
    # test_permutation_compute_sigmas_only_left_wires: variables zero, two .. nine = ids 0 .. 8
    dict(vars=[[0, 0, 0, 0], [0, 1, 2, 3], [4, 5, 6, 7], [8, 8, 8, 8]], n_vars=9,
         sigma=[[wd(R, 0), wd(L, 2), wd(L, 3), wd(L, 0)], [wd(L, 1), wd(R, 1), wd(R, 2), wd(R, 3)],
                [wd(O, 0), wd(O, 1), wd(O, 2), wd(O, 3)], [wd(F, 1), wd(F, 2), wd(F, 3), wd(F, 0)]]),
    # test_permutation_compute_sigmas: variables one .. four = ids 0 .. 3
    dict(vars=[[0, 1, 2, 1], [0, 0, 2, 0], [1, 1, 0, 2], [3, 3, 3, 3]], n_vars=4,
         sigma=[[wd(R, 0), wd(O, 1), wd(R, 2), wd(O, 0)], [wd(R, 1), wd(O, 2), wd(O, 3), wd(L, 0)],
                [wd(L, 1), wd(L, 3), wd(R, 3), wd(L, 2)], [wd(F, 1), wd(F, 2), wd(F, 3), wd(F, 0)]]),
]


def _encode(oracle, code):
    """K_wire * omega_4^gate as the reference test writes it (K = 1, 7, 13, 17; w = group_gen of the size-4 domain)."""
    w, winv, ninv = (np.zeros(4, np.uint64) for _ in range(3))
    oracle.lib.zpo_fr_root_of_unity(2, _p(w), _p(winv), _p(ninv))
    k = fr_small(oracle, [1, 7, 13, 17])
    pw = [fr_small(oracle, [1])[0]]
    for _ in range(3):
        pw.append(oracle.fr_op(2, pw[-1].reshape(1, 4), w.reshape(1, 4))[0])
    return oracle.fr_op(2, k[code >> 28].reshape(1, 4), pw[code & 0xfffffff].reshape(1, 4))[0]


@pytest.mark.parametrize("case", SIGMA_CASES)
def test_oracle_sigma_reference_vectors(oracle, case):
    v = np.array(case["vars"], dtype=np.uint32)
    code = np.zeros(16, dtype=np.uint32)
    enc = np.zeros((16, 4), dtype=np.uint64)
    oracle.lib.zpo_sigma_from_wires.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                oracle_lib.u64p]
    oracle.lib.zpo_sigma_from_wires(2, 4, v.ctypes.data, case["n_vars"], code.ctypes.data, _p(enc))
    want = np.array(case["sigma"], dtype=np.uint32).reshape(16)
    assert np.array_equal(code, want)
    for i in range(16):
        assert np.array_equal(enc[i], _encode(oracle, int(want[i])))


# ---- util.rs:303-321
def test_lc_matches_power_sum(oracle):
    oracle.lib.zpo_lc.argtypes = [ctypes.c_size_t, oracle_lib.u64p, oracle_lib.u64p, oracle_lib.u64p]
    for it in range(10):
        v = oracle.random_fr(100 + it, 5)
        ch = oracle.random_fr(200 + it, 1)
        out = np.zeros(4, dtype=np.uint64)
        oracle.lib.zpo_lc(5, _p(v), _p(ch), _p(out))
        expected, p = v[0:1].copy(), ch.copy()
        for k in range(1, 5):
            expected = oracle.fr_op(0, expected, oracle.fr_op(2, v[k:k + 1].copy(), p))
            p = oracle.fr_op(2, p, ch)
        assert np.array_equal(out, expected[0])


def _product_compress(pkg, lib, oracle):
    ctx = pkg.ProverContext(8, lib)
    oracle.lib.zpo_lc.argtypes = [ctypes.c_size_t, oracle_lib.u64p, oracle_lib.u64p, oracle_lib.u64p]
    cols = [oracle.random_fr(300 + k, 257) for k in range(4)]
    ch = oracle.random_fr(7, 1)[0]
    out = ctx.multiset_compress(cols, ch)
    for i in [0, 1, 100, 256]:
        v = np.stack([c[i] for c in cols])
        want = np.zeros(4, dtype=np.uint64)
        oracle.lib.zpo_lc(4, _p(v), _p(ch), _p(want))
        assert np.array_equal(out[i], want)
    ctx.close()


def test_emulated_multiset_compress(pkg, emu_lib, oracle):
    _product_compress(pkg, emu_lib, oracle)


@pytest.mark.gpu
def test_gpu_multiset_compress(pkg, gpu_lib, oracle):
    _product_compress(pkg, gpu_lib, oracle)
