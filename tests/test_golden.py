"""Known-answer tests against the committed fixtures of tests/golden/ (see make_golden.py for their provenance)."""
import os

import numpy as np
import pytest

import oracle_lib

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_golden(oracle):
    oc = oracle_lib.OracleCircuit(oracle, 4, 42, 7, 0)
    assert np.array_equal(oc.prove()[0], np.load(os.path.join(G, "proof_height4_w42_tau7.npy")))
    oc.close()
    d = np.load(os.path.join(G, "ntt_2e6_seed1.npz"))
    assert np.array_equal(oracle.random_fr(1, 64), d["x"])
    for kind, key in enumerate(["fft", "ifft", "coset_fft", "coset_ifft"]):
        assert np.array_equal(oracle.ntt(kind, d["x"]), d[key])
    m = np.load(os.path.join(G, "msm_256_tau7_seed2.npz"))
    assert np.array_equal(oracle.msm(m["points"], m["scalars"]), m["result"])
    ch = oracle.transcript_script(b"Merkle tree", [("append", b"pi", bytes(range(48))), ("challenge", b"zeta", 31),
                                                   ("append", b"zeta", bytes(32)), ("challenge", b"beta", 31)])
    assert ch.hex() == open(os.path.join(G, "transcript_challenges.hex")).read().strip()


def _product_vs_golden(pkg, lib, oracle):
    ctx = pkg.ProverContext(8, lib)
    d = np.load(os.path.join(G, "ntt_2e6_seed1.npz"))
    for kind, key in enumerate(["fft", "ifft", "coset_fft", "coset_ifft"]):
        assert np.array_equal(ctx.ntt(kind, d["x"]), d[key])
    m = np.load(os.path.join(G, "msm_256_tau7_seed2.npz"))
    assert np.array_equal(ctx.msm_points(m["points"], m["scalars"]), m["result"])
    ctx.close()
    for height, n_lookup, name in [(3, 12, "proof_height3_lookup12_w42_tau7.npy"), (4, 0, "proof_height4_w42_tau7.npy")]:
        oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, n_lookup, with_pk=False)
        c = pkg.ProverContext(oc.log_n, lib)
        c.load_srs(oc.srs())
        c.preprocess(oc.selector_evals(), oc.tables())
        circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
        assert np.array_equal(c.prove(circ).to_words(), np.load(os.path.join(G, name)))
        c.close()
        oc.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden(pkg, gpu_lib, oracle):
    _product_vs_golden(pkg, gpu_lib, oracle)
