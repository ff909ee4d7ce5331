"""End-to-end run at a given Merkle height on the GPU: circuit synthesis (CPU oracle front end), SRS generation
and preprocessing on the device, proofs with timing, verification with the oracle's verifier restatement."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=15)
    ap.add_argument("--proofs", type=int, default=3)
    ap.add_argument("--verify", type=int, default=1)
    ap.add_argument("--microbench", type=int, default=1)
    args = ap.parse_args()
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    t = time.time()
    oc = oracle_lib.OracleCircuit(orc, args.height, 42, 7, 0, with_pk=False, with_srs=False)
    print("circuit synthesis %.1fs cs.n=%d logN=%d" % (time.time() - t, oc.cs_n, oc.log_n), flush=True)
    ctx = pkg.ProverContext(oc.log_n, lib)
    if args.microbench:
        for mode, name in [(0, "IMAD G/s"), (1, "IMAD.WIDE G/s"), (2, "Fq mul G/s")]:
            print("int pipe", name, "%.1f" % ctx.bench_int_pipe(mode), flush=True)
    t = time.time()
    ctx.generate_srs(oc.tau())
    print("device SRS %.2fs" % (time.time() - t), flush=True)
    t = time.time()
    sel = oc.selector_evals()
    ctx.preprocess(sel, oc.tables())
    del sel
    print("device preprocess %.2fs" % (time.time() - t), flush=True)
    wires, ql, pi = oc.wires(), oc.q_lookup(), oc.pi_canonical()
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, ql, pi, *wires)
    proof = None
    for i in range(args.proofs):
        t = time.time()
        p = ctx.prove(circ)
        dt = time.time() - t
        print("proof %d wall %.3fs timing %s" % (i, dt, json.dumps(ctx.last_timing())), flush=True)
        w = p.to_words()
        if proof is not None:
            assert np.array_equal(w, proof), "proof not deterministic"
        proof = w
    if args.verify:
        t = time.time()
        oc.set_vk(ctx.verifier_key())
        print("device verifier key %.2fs" % (time.time() - t), flush=True)
        t = time.time()
        ok, detail = oc.verify(proof)
        print("verifier restatement: ok=%s detail=%d (%.2fs)" % (ok, detail, time.time() - t), flush=True)
        assert ok
    # NTT / MSM sweeps (device-resident)
    n = oc.n
    ctx.bench_alloc(0, 8 * n)
    ctx.bench_alloc(1, 8 * n)
    x = orc.random_fr(1, n)
    ctx.bench_upload(0, x)
    for logn, kind in [(oc.log_n, 0), (oc.log_n, 1), (oc.log_n + 3, 2), (oc.log_n + 3, 3)]:
        ms = ctx.bench_ntt(kind, logn, 0, 1, 5)
        print("ntt kind=%d logn=%d %.3f ms  (%.1f GB/s algorithmic)" % (kind, logn, ms, 64.0 * (1 << logn) / ms / 1e6), flush=True)
    ms, out, bd = ctx.bench_msm(0, n, 3)
    print("msm n=2^%d %.2f ms breakdown %s" % (oc.log_n, ms, json.dumps(bd)), flush=True)


if __name__ == "__main__":
    main()
