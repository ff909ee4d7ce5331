"""Runs the REFERENCE's own native prover (PNP lib/ compiled unmodified for sm_100 into oracle/_ref/libzprize_ref.so
by oracle/build_pnp_ref.sh) on an oracle-generated Merkle circuit through its `gen_proof` FFI symbol and stores the
ProofC image.  Used by tests/test_gpu_vs_pnp_reference.py in a subprocess (the reference exits the process on errors
and prints debug output)."""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=4)
    ap.add_argument("--out", required=True)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--lib", default="libzprize_ref.so",
                    help="libzprize_ref.so (unmodified; dies above HEIGHT=4 on this box) or libzprize_ref_patched.so (the reference's double "
                         "destruction and 18-byte MSM result buffer fixed at build time, see oracle/build_pnp_ref.sh)")
    args = ap.parse_args()
    pkg = load_package()
    trace = os.path.join(ROOT, "oracle", "libsegv_trace.so")
    if os.path.exists(trace) and not os.environ.get("ZP_NO_SEGV_TRACE"):
        ctypes.CDLL(trace)  # native backtrace on SIGSEGV: the UNMODIFIED reference crashes above HEIGHT=4
    path = os.path.join(ROOT, "oracle", "_ref", args.lib)
    ref = ctypes.CDLL(path)
    ref.gen_proof.restype = pkg.ProofC
    ref.gen_proof.argtypes = [pkg.CircuitC, pkg.ProverKeyC, pkg.CommitKeyC]
    orc = oracle_lib.load()
    oc = oracle_lib.OracleCircuit(orc, args.height, 42, 7, 0)
    names = pkg.PK_POLY_NAMES + pkg.PK_SIGMA_NAMES
    co, ev, tb = oc.pk_coeffs(), oc.pk_evals(), oc.tables()
    le, vh = oc.linear_evaluations(), oc.v_h_coset_8n()
    pk = pkg.make_prover_key(dict(zip(names, co)), dict(zip(names, ev)), tb, le, vh)
    srs = oc.srs()
    gamma = np.zeros((2, 12), dtype=np.uint64)
    ck = pkg.CommitKeyC()
    ck.powers_of_g = pkg.as_u64p(srs)
    ck.powers_of_gamma_g = pkg.as_u64p(gamma)
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    import time
    for i in range(args.repeat):
        t = time.time()
        proof = ref.gen_proof(circ, pk, ck)
        print("reference gen_proof call %d: %.3f s" % (i, time.time() - t), flush=True)
    np.save(args.out, proof.to_words())
    pinned = os.path.join(ROOT, "tests", "golden", "proof_height%d_w42_tau7.npy" % args.height)
    if os.path.exists(pinned):  # large sizes: the oracle's proof was pinned once (tests/golden/make_golden_large.py)
        oracle_proof = np.load(pinned)
    else:
        oracle_proof, _ = oc.prove()
    np.save(args.out.replace(".npy", "_oracle.npy"), oracle_proof)
    same = bool(np.array_equal(proof.to_words(), oracle_proof))
    print("reference proof written; equals the oracle's proof: %s" % same, flush=True)


if __name__ == "__main__":
    main()
