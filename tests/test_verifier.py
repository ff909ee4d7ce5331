"""Product-side verifier with REAL pairings (csrc/verifier.cu, csrc/pairing.hpp; host code — runs without a GPU, here
through the g++ build of the same sources and, when nvcc's library is present, through libzprize_b200.so itself).

Pins: (1) the cube of our GT element of e(aG, bH) equals, byte for byte, what the reference's own vendored blst computes
with `blst_miller_loop` + `blst_final_exp` (oracle/_ref/libref_blst.so; blst's hard part carries the customary factor 3); (2) bilinearity; (3) the oracle's proofs — Merkle,
lookup and gadget circuits — are accepted, tampered ones rejected with the failing opening reported; (4) batch
verification (`zp_proof_verify_batch`) accepts honest batches and rejects a batch with one bad proof.
Reference: "Prize 1B/plonk-core/src/proof_system/proof.rs":123-443, verifier.rs:106-125."""
import ctypes

import numpy as np
import pytest

import oracle_lib
from oracle_lib import _p

ONE = np.array([8589934590, 6378425256633387010, 11064306276430008309, 1739710354780652911], dtype=np.uint64)


def g1_gen(oracle):
    g = np.zeros(12, dtype=np.uint64)
    oracle.lib.zpo_g1_generator(_p(g))
    return g


def test_pairing_matches_reference_blst(pkg, emu_lib, oracle):
    blst = oracle_lib.ref_lib("libref_blst.so")
    if blst is None:
        pytest.skip("oracle/_ref/libref_blst.so not built")
    g = g1_gen(oracle)
    sc = oracle.random_fr(77, 4)
    for i in range(2):
        P = oracle.g1_mul(g, sc[2 * i])
        Q = pkg.g2_mul_generator(sc[2 * i + 1], emu_lib)
        ours, is_one = pkg.pairing_product(P, Q, emu_lib)
        assert not is_one
        ml, fe = np.zeros(72, dtype=np.uint64), np.zeros(72, dtype=np.uint64)
        blst.blst_miller_loop(_p(ml), _p(Q), _p(P))
        blst.blst_final_exp(_p(fe), _p(ml))
        # blst's final exponentiation uses the usual x-chain for the hard part, which raises to 3 (q^4 - q^2 + 1) / r;
        # ours raises to exactly (q^12 - 1) / r.  3 is coprime to r, so both are the same pairing up to that fixed power.
        sq, cube = np.zeros(72, dtype=np.uint64), np.zeros(72, dtype=np.uint64)
        blst.blst_fp12_mul(_p(sq), _p(ours.copy()), _p(ours.copy()))
        blst.blst_fp12_mul(_p(cube), _p(sq), _p(ours.copy()))
        assert np.array_equal(cube, fe), "GT element^3 differs from blst_final_exp(blst_miller_loop(Q, P))"
    # G2 generator and scalar multiplication agree with blst as well
    blst.blst_p2_affine_generator.restype = oracle_lib.u64p
    h_ref = np.ctypeslib.as_array(blst.blst_p2_affine_generator(), shape=(24,)).copy()
    assert np.array_equal(pkg.g2_mul_generator(ONE, emu_lib), h_ref)


def test_pairing_bilinearity(pkg, emu_lib, oracle):
    g = g1_gen(oracle)
    a = oracle.random_fr(5, 1)[0]
    neg_one = oracle.fr_op(6, ONE.reshape(1, 4))[0]
    aG, mG = oracle.g1_mul(g, a), oracle.g1_mul(g, neg_one)
    H, aH = pkg.g2_mul_generator(ONE, emu_lib), pkg.g2_mul_generator(a, emu_lib)
    _, is_one = pkg.pairing_product(np.stack([aG, mG]), np.stack([H, aH]), emu_lib)  # e(aG, H) e(-G, aH) = 1
    assert is_one
    _, is_one = pkg.pairing_product(np.stack([aG, mG]), np.stack([H, H]), emu_lib)
    assert not is_one


def _verifier_for(pkg, lib, oracle, oc):
    """Verifier key = commitments to the 19 prover-key polynomials + 4 table polynomials, here [p(tau)] G with the
    known trapdoor; beta_h = tau H."""
    co = oc.pk_coeffs()
    comms = [oc.commit_with_tau(c) for c in co] + [oc.commit_with_tau(oracle.ntt(1, t)) for t in oc.tables()]
    return pkg.Verifier(oc.n, np.stack(comms), pkg.g2_mul_generator(oc.tau(), lib), lib)


def _pi_mont(oracle, oc):
    pi = oc.pi_canonical()
    if not pi.any():
        return None
    return oracle.fr_op(5, pi.reshape(1, 4))[0]


@pytest.mark.parametrize("height,kind,n_lookup", [(3, 0, 0), (3, 0, 12), (0, 1, 12), (0, 3, 0)])
def test_verifier_accepts_oracle_proofs_and_rejects_tampering(pkg, emu_lib, oracle, height, kind, n_lookup):
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, n_lookup, kind=kind)
    proof, _ = oc.prove()
    v = _verifier_for(pkg, emu_lib, oracle, oc)
    pim = _pi_mont(oracle, oc)
    assert v.verify(proof, oc.pi_pos, pim) == (True, 3)
    assert oc.verify(proof)[0]  # the oracle's trapdoor verifier agrees
    bad = proof.copy()
    bad[12 * 19 + 4 * 7] ^= 1  # permutation_eval: only the shifted opening and r0 see it
    ok, detail = v.verify(bad, oc.pi_pos, pim)
    assert not ok
    bad = proof.copy()
    bad[12 * 19] ^= 1  # a_eval
    assert not v.verify(bad, oc.pi_pos, pim)[0]
    if pim is not None:
        wrong_pi = oracle.fr_op(0, pim.reshape(1, 4), ONE.reshape(1, 4))[0]
        assert not v.verify(proof, oc.pi_pos, wrong_pi)[0]
    # swapped openings
    bad = proof.copy()
    bad[12 * 17:12 * 18], bad[12 * 18:12 * 19] = proof[12 * 18:12 * 19], proof[12 * 17:12 * 18]
    assert v.verify(bad, oc.pi_pos, pim) == (False, 0)
    v.close()
    oc.close()


def test_batch_verification(pkg, emu_lib, oracle):
    """Same circuit (same verifier key), three different witnesses."""
    ocs = [oracle_lib.OracleCircuit(oracle, 3, seed, 7, 0) for seed in (42, 43, 44)]
    proofs = [oc.prove()[0] for oc in ocs]
    assert not np.array_equal(proofs[0], proofs[1])
    v = _verifier_for(pkg, emu_lib, oracle, ocs[0])
    pos = [oc.pi_pos for oc in ocs]
    vals = np.stack([_pi_mont(oracle, oc) for oc in ocs])
    for p, oc, val in zip(proofs, ocs, vals):
        assert v.verify(p, oc.pi_pos, val)[0]
    assert v.verify_batch(proofs, pos, vals)
    bad = [p.copy() for p in proofs]
    bad[1][12 * 19 + 4] ^= 1
    assert not v.verify_batch(bad, pos, vals)
    assert not v.verify_batch(proofs, pos, vals[::-1].copy())
    v.close()
    for oc in ocs:
        oc.close()


def test_verifier_in_the_shipped_library(pkg, oracle):
    """The same entry points in libzprize_b200.so (nvcc build) — host code, no device needed."""
    import os
    if not os.path.exists(pkg.LIB_PATH):
        pytest.skip("libzprize_b200.so not built")
    lib = pkg.load_library()
    oc = oracle_lib.OracleCircuit(oracle, 3, 42, 7, 0)
    proof, _ = oc.prove()
    v = _verifier_for(pkg, lib, oracle, oc)
    assert v.verify(proof, oc.pi_pos, _pi_mont(oracle, oc)) == (True, 3)
    v.close()
    oc.close()


@pytest.mark.gpu
def test_device_proof_and_device_verifier_key_verify_with_pairings(pkg, gpu_lib, oracle):
    """End to end on the box: device preprocessing -> verifier key (device MSMs), device proof, pairing verifier."""
    oc = oracle_lib.OracleCircuit(oracle, 8, 42, 7, 0, with_pk=False)
    ctx = pkg.ProverContext(oc.log_n, gpu_lib)
    ctx.load_srs(oc.srs())
    v_, c_, nv = oc.wiring()
    ctx.preprocess_wiring(oc.selector_evals()[:15], v_, c_, nv, oc.tables())
    vk = ctx.verifier_key()
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = ctx.prove(circ)
    v = pkg.Verifier(oc.n, vk, pkg.g2_mul_generator(oc.tau(), gpu_lib), gpu_lib)
    assert v.verify(proof, oc.pi_pos, _pi_mont(oracle, oc)) == (True, 3)
    words = proof.to_words()
    words[12 * 4] ^= 1  # z commitment x-coordinate: not on the curve any more -> error, not acceptance
    with pytest.raises(pkg.ZprizeError):
        v.verify(words, oc.pi_pos, _pi_mont(oracle, oc))
    v.close()
    ctx.close()
    oc.close()
