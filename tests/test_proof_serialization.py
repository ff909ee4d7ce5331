"""ark-serialize wire format of the proof: product serializer vs the oracle's, layout facts, round trip."""
import ctypes

import numpy as np

import oracle_lib


def test_proof_serialize_matches_oracle_and_round_trips(pkg, emu_lib, oracle):
    oc = oracle_lib.OracleCircuit(oracle, 4, 42, 7, 0)
    words, _ = oc.prove()
    oc.close()
    proof = pkg.ProofC.from_buffer_copy(words.tobytes())
    data = pkg.proof_serialize(proof, emu_lib)
    assert len(data) == pkg.PROOF_SERIALIZED_BYTES == 17 * 48 + 2 * 49 + 16 * 32 + 8 + sum(8 + len(n) + 32 for n in [
        "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval", "q_hr_eval", "q_h4_eval", "a_next_eval", "b_next_eval",
        "d_next_eval"])
    out = ctypes.create_string_buffer(2048)
    oracle.lib.zpo_proof_serialize.restype = ctypes.c_size_t
    oracle.lib.zpo_proof_serialize.argtypes = [oracle_lib.u64p, ctypes.c_void_p]
    n = oracle.lib.zpo_proof_serialize(oracle_lib._p(words), out)
    assert out.raw[:n] == data
    # f, h_1, h_2, t_7, t_8 are the identity: 47 zero bytes and the infinity flag (bit 6)
    for idx in [5, 6, 7, 15, 16]:
        assert data[48 * idx:48 * idx + 48] == b"\x00" * 47 + b"\x40"
    assert data[17 * 48 + 48] == 0 and data[17 * 48 + 49 + 48] == 0  # random_v = None
    back = pkg.proof_deserialize(data, emu_lib)
    assert np.array_equal(back.to_words(), words)
    bad = bytearray(data)
    bad[3] ^= 1  # x no longer on the curve (with overwhelming probability) or a different point
    try:
        other = pkg.proof_deserialize(bytes(bad), emu_lib)
        assert not np.array_equal(other.to_words(), words)
    except pkg.ZprizeError:
        pass
