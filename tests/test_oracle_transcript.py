"""Merlin / STROBE known-answer test and ark-serialize encodings of the oracle transcript."""
import ctypes

import numpy as np

import oracle_lib
from oracle_lib import _p


def test_merlin_equivalence_simple_vector(oracle):
    # merlin 3.0.0 tests::equivalence_simple (SURVEY §8c pin (6))
    out = ctypes.create_string_buffer(32)
    oracle.lib.zpo_transcript_kat(b"test protocol", b"some label", b"some data", 9, b"challenge", out, 32)
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_long_messages_cross_rate_boundary(oracle):
    # absorbing more than the 166-byte STROBE rate must stay self-consistent between split and joined messages
    a = oracle.transcript_script(b"p", [("append", b"l", b"x" * 1000), ("challenge", b"c", 64)])
    b = oracle.transcript_script(b"p", [("append", b"l", b"x" * 1000), ("challenge", b"c", 64)])
    c = oracle.transcript_script(b"p", [("append", b"l", b"x" * 999 + b"y"), ("challenge", b"c", 64)])
    assert a == b and a != c and len(a) == 64


def test_g1_compressed_encoding(oracle):
    g = np.zeros(12, dtype=np.uint64)
    oracle.lib.zpo_g1_generator(_p(g))
    out = ctypes.create_string_buffer(48)
    oracle.lib.zpo_g1_serialize(_p(g), out)
    b = out.raw
    # x of the generator, little-endian, top byte 0x17; generator's y is the lexicographically smaller root -> no sign bit
    x_be = bytes.fromhex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
    assert b[:47] == x_be[::-1][:47]
    assert b[47] & 0x3f == 0x17
    y = int("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1", 16)
    q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab", 16)
    assert bool(b[47] & 0x80) == (y > q - y)
    # infinity: FFI encoding (0, Mont(1)) -> all-zero x with bit 6 set
    inf = np.zeros(12, dtype=np.uint64)
    inf[6:] = [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d,
               0x15f65ec3fa80e493]
    oracle.lib.zpo_g1_serialize(_p(inf), out)
    assert out.raw == b"\x00" * 47 + b"\x40"
