"""Memory-safety check of the product's kernels WITHOUT a GPU: the .cu sources compiled against the CPU emulation layer
(tests/emu) with -fsanitize=address, so every out-of-bounds access to a "device" buffer (malloc under emulation) or
use-after-free aborts with a report.  Exercises the MSM pipeline (batch-affine rounds with the x[] / y[] partial sums, slot
table with long runs, work-list fold, host-side top of the inversion tree, window table with the shared inversion, batches)
and whole proofs (lookups, second-stream NTT path).  Driven by tools/asan_emu/run.sh.  TEST INFRASTRUCTURE (uses oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    rounds = sys.argv[2] if len(sys.argv) > 2 else "2"
    os.environ["ZP_MSM_PRECOMP_MIN_LOG"] = "8"
    os.environ["ZP_MSM_BA_ROUNDS"] = rounds
    os.environ["ZP_MSM_BA_MIN_LOG"] = "4"
    pkg = load_package()
    lib = pkg.load_library(sys.argv[1])
    orc = oracle_lib.load()
    # operator MSM over caller points: uniform, skewed (500 of 700 scalars equal: runs of 250 pairs, long fold list), repeated points
    ctx = pkg.ProverContext(8, lib)
    n = 700
    pts, _ = orc.srs(7, n)
    sc = orc.random_fr(2, n)
    assert np.array_equal(ctx.msm_points(pts, sc, 6), orc.msm(pts, sc))
    three = np.zeros((1, 4), dtype=np.uint64)
    three[0, 0] = 3
    sk = sc.copy()
    sk[100:600] = orc.fr_op(5, three)[0]
    assert np.array_equal(ctx.msm_points(pts, sk, 6), orc.msm(pts, sk))
    p2, s2 = pts.copy(), sc.copy()
    p2[10:40] = p2[10]
    s2[10:40] = s2[10]
    assert np.array_equal(ctx.msm_points(p2, s2, 6), orc.msm(p2, s2))
    print("rounds=%s: operator MSMs (uniform / skewed / repeated points) ok" % rounds, flush=True)
    ctx.close()
    # whole proofs: precomputed window table, batched commitments, lookups; then the second-stream NTT path
    for overlap in ("0", "1"):
        os.environ["ZP_NTT_OVERLAP"] = overlap
        oc = oracle_lib.OracleCircuit(orc, 3, 42, 7, 12)
        ref, _ = oc.prove()
        c = pkg.ProverContext(oc.log_n, lib)
        c.load_srs(oc.srs())
        c.preprocess(oc.selector_evals(), oc.tables())
        circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
        assert np.array_equal(c.prove(circ).to_words(), ref)
        scal = [orc.random_fr(20 + k, oc.n) for k in range(3)]
        outs = c.msm_batch(scal)
        for k in range(3):
            assert np.array_equal(outs[k], orc.msm(oc.srs(), scal[k]))
        print("rounds=%s overlap=%s: proof with lookups + 3-member MSM batch over the window table ok" % (rounds, overlap), flush=True)
        c.close()
        oc.close()


if __name__ == "__main__":
    main()
