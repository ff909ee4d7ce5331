"""Whole-proof byte parity at the sizes where gen_proof takes its PRODUCTION route — precomputed MSM window tables,
batch-affine bucket rounds, batched commitments, 3-pass NTTs (N >= 2^16): HEIGHT=10 (N = 2^17), 12 (2^19) and the
benchmark size HEIGHT=15 (2^22).  The expected bytes are the CPU oracle's proofs pinned under tests/golden/ by
tests/golden/make_golden_large.py (the oracle needs 15 s ... 10 min per proof, so it is not re-run here); the
circuit front end (witness + selectors) is rebuilt from the same seeds."""
import os

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _prove(pkg, lib, oracle, height):
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, 0, with_pk=False, with_srs=False)
    ctx = pkg.ProverContext(oc.log_n, lib)
    ctx.generate_srs(oc.tau())
    ctx.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = ctx.prove(circ).to_words()
    again = ctx.prove(circ).to_words()
    ctx.close()
    oc.close()
    return proof, again


@pytest.mark.parametrize("height", [10, 12, 15])
def test_gen_proof_equals_pinned_oracle_proof(pkg, gpu_lib, oracle, height):
    path = os.path.join(G, "proof_height%d_w42_tau7.npy" % height)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % os.path.basename(path))
    proof, again = _prove(pkg, gpu_lib, oracle, height)
    assert np.array_equal(proof, again), "proof differs between two runs on the same context"
    assert np.array_equal(proof, np.load(path)), "device proof differs from the pinned oracle proof"


def test_gen_proof_with_second_stream_ntts(pkg, gpu_lib, oracle, monkeypatch):
    """ZP_NTT_OVERLAP=1 (experiment, off by default): the coset NTTs of the wires and of z(X) run on the prover's second stream
    concurrently with the commitment MSMs and are joined by events before the quotient pass — same proof bytes."""
    path = os.path.join(G, "proof_height12_w42_tau7.npy")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    monkeypatch.setenv("ZP_NTT_OVERLAP", "1")  # read when the context is created
    proof, again = _prove(pkg, gpu_lib, oracle, 12)
    assert np.array_equal(proof, again)
    assert np.array_equal(proof, np.load(path))


def test_lookup_proof_through_production_routes(pkg, gpu_lib, oracle):
    """HEIGHT=9 Merkle circuit + 2000 plookup rows (N = 2^16): compress / query table / combine_split / z2 and the lookup
    terms of the quotient and linearisation at a size where the MSM and NTT take their production routes; expected bytes
    from the oracle run here (a few seconds), plus acceptance by the product's pairing verifier."""
    oc = oracle_lib.OracleCircuit(oracle, 9, 42, 7, 2000)
    ref, _ = oc.prove()
    ctx = pkg.ProverContext(oc.log_n, gpu_lib)
    ctx.load_srs(oc.srs())
    v, c, nv = oc.wiring()
    ctx.preprocess_wiring(oc.selector_evals()[:15], v, c, nv, oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = ctx.prove(circ).to_words()
    assert np.array_equal(proof, ref)
    ver = pkg.Verifier(oc.n, ctx.verifier_key(), pkg.g2_mul_generator(oc.tau(), gpu_lib), gpu_lib)
    assert ver.verify(proof, oc.pi_pos, oracle.fr_op(5, oc.pi_canonical().reshape(1, 4))[0]) == (True, 3)
    ver.close()
    ctx.close()
    oc.close()
