"""Pins the oracle's derived constants against the literal constants the reference carries in-tree
(SURVEY §8c "the only fixed numbers in-tree are field/curve constants")."""
import ctypes

import numpy as np

import oracle_lib
from oracle_lib import _p


def _arr(n):
    return np.zeros(n, dtype=np.uint64)


def test_fr_constants_match_reference_literals(oracle):
    m, one, rr, tw, g = (_arr(4) for _ in range(5))
    inv = _arr(1)
    oracle.lib.zpo_fr_constants(_p(m), _p(one), _p(rr), _p(inv), _p(tw), _p(g))
    # "Prize 1B/plonk-core/lib/PLONK/src/bls12_381/fr.cuh":23-53
    assert list(one) == [8589934590, 6378425256633387010, 11064306276430008309, 1739710354780652911]
    assert list(m) == [18446744069414584321, 6034159408538082302, 3691218898639771653, 8353516859464449352]
    assert list(tw) == [13381757501831005802, 6564924994866501612, 789602057691799140, 6625830629041353339]
    assert list(g) == [64424509425, 1721329240476523535, 18418692815241631664, 3824455624000121028]
    # "Prize 1B/plonk-core/lib/PLONK/utils/mont/cpu/ff/bls12-381.hpp":12-27
    assert list(rr) == [0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]
    assert int(inv[0]) == 0xfffffffeffffffff


def test_fq_constants_match_reference_literals(oracle):
    m, one, rr = (_arr(6) for _ in range(3))
    inv = _arr(1)
    oracle.lib.zpo_fq_constants(_p(m), _p(one), _p(rr), _p(inv))
    # "Prize 1B/plonk-core/lib/PLONK/utils/mont/cpu/ff/bls12-381.hpp":34-60
    assert list(m) == [0xb9feffffffffaaab, 0x1eabfffeb153ffff, 0x6730d2a0f6b0f624, 0x64774b84f38512bf,
                       0x4b1ba7b6434bacd7, 0x1a0111ea397fe69a]
    assert list(rr) == [0xf4df1f341c341746, 0x0a76e6a609d104f1, 0x8de5476c4c95b6d5, 0x67eb88a9939d83c0,
                        0x9a793e85b519952d, 0x11988fe592cae3aa]
    assert list(one) == [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745,
                         0x5c071a97a256ec6d, 0x15f65ec3fa80e493]
    assert int(inv[0]) == 0x89f3fffcfffcfffd


def test_jubjub_constants_match_reference_literals(oracle):
    a, d = _arr(4), _arr(4)
    oracle.lib.zpo_jubjub(_p(a), _p(d))
    # "Prize 1B/plonk-core/lib/PLONK/src/bls12_381/edwards.cu":5-31
    assert list(a) == [18446744060824649731, 18102478225614246908, 11073656695919314959, 6613806504683796440]
    assert list(d) == [3049539848285517488, 18189135023605205683, 8793554888777148625, 6339087681201251886]


def test_generator_on_curve(oracle):
    g = _arr(12)
    oracle.lib.zpo_g1_generator(_p(g))
    assert oracle.lib.zpo_g1_on_curve(_p(g)) == 1


def test_roots_of_unity_chain(oracle):
    # omega_{2^k}^2 == omega_{2^(k-1)}, omega_{2^k}^(2^k) == 1, omega * omega^-1 == 1, n * n^-1 == 1
    one = np.array([[8589934590, 6378425256633387010, 11064306276430008309, 1739710354780652911]], dtype=np.uint64)
    prev = None
    for k in range(1, 27):
        w, wi, ni = _arr(4), _arr(4), _arr(4)
        oracle.lib.zpo_fr_root_of_unity(k, _p(w), _p(wi), _p(ni))
        assert np.array_equal(oracle.fr_op(2, w.reshape(1, 4), wi.reshape(1, 4)), one)
        if prev is not None:
            assert np.array_equal(oracle.fr_op(2, w.reshape(1, 4), w.reshape(1, 4))[0], prev)
        prev = w
