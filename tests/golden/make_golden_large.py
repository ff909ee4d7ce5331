"""Pins the CPU oracle's proof at the benchmark sizes (run once here, ~10-20 min and ~45 GB of host RAM at HEIGHT=15):
  python tests/golden/make_golden_large.py 15 12 10
writes tests/golden/proof_height<H>_w42_tau7.npy (the 2656-byte ProofC image as 332 u64 words) + its SHA-256.
The fixtures are ORACLE outputs (same provenance as make_golden.py): bench.py and the -m gpu tests compare the device
proof with them byte for byte, so HEIGHT=15 parity no longer rests on the verifier restatement alone."""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402

if __name__ == "__main__":
    orc = oracle_lib.load()
    for h in [int(x) for x in sys.argv[1:]]:
        t0 = time.time()
        oc = oracle_lib.OracleCircuit(orc, h, 42, 7, 0)
        t1 = time.time()
        proof, secs = oc.prove()
        ok, _ = oc.verify(proof)
        assert ok
        np.save(os.path.join(HERE, "proof_height%d_w42_tau7.npy" % h), proof)
        sha = hashlib.sha256(proof.tobytes()).hexdigest()
        with open(os.path.join(HERE, "proof_height%d_w42_tau7.sha256" % h), "w") as f:
            f.write(sha + "\n")
        print("HEIGHT=%d cs.n=%d N=2^%d setup %.1fs prove %.1fs threads %d sha256 %s" %
              (h, oc.cs_n, oc.log_n, t1 - t0, secs, orc.lib.zpo_num_threads(), sha), flush=True)
        oc.close()
