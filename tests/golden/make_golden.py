"""Generates the golden fixtures of tests/golden/.

The reference repository holds NO golden vectors for this path (no proof bytes, commitments, transcripts
or NTT vectors anywhere; native code has no tests) and its Rust prover cannot be run in this image, so these
fixtures are produced by the CPU oracle (oracle/) — after the oracle itself has been pinned against the reference's
constants, vendored blst, strobe.cpp and the Merlin KAT (tests/test_oracle_*.py).  They are regression pins for the
oracle and known-answer inputs/outputs for the GPU path, not independent evidence about the Rust prover.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402


def main():
    orc = oracle_lib.load()
    oc = oracle_lib.OracleCircuit(orc, 4, 42, 7, 0)
    proof, _ = oc.prove()
    assert oc.verify(proof)[0]
    np.save(os.path.join(HERE, "proof_height4_w42_tau7.npy"), proof)
    oc.close()
    oc = oracle_lib.OracleCircuit(orc, 3, 42, 7, 12)
    proof, _ = oc.prove()
    assert oc.verify(proof)[0]
    np.save(os.path.join(HERE, "proof_height3_lookup12_w42_tau7.npy"), proof)
    oc.close()
    x = orc.random_fr(1, 64)
    np.savez(os.path.join(HERE, "ntt_2e6_seed1.npz"), x=x, fft=orc.ntt(0, x), ifft=orc.ntt(1, x), coset_fft=orc.ntt(2, x),
             coset_ifft=orc.ntt(3, x))
    pts, tau = orc.srs(7, 256)
    sc = orc.random_fr(2, 256)
    np.savez(os.path.join(HERE, "msm_256_tau7_seed2.npz"), points=pts, scalars=sc, tau=tau, result=orc.msm(pts, sc))
    ch = orc.transcript_script(b"Merkle tree", [("append", b"pi", bytes(range(48))), ("challenge", b"zeta", 31),
                                                ("append", b"zeta", bytes(32)), ("challenge", b"beta", 31)])
    open(os.path.join(HERE, "transcript_challenges.hex"), "w").write(ch.hex() + "\n")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
