"""Cross-run against the REFERENCE ITSELF (SURVEY §8c pin (5)): PNP's native prover, compiled unmodified for sm_100
into oracle/_ref/libzprize_ref.so, is executed on the GPU box on a Merkle-shaped circuit; its ProofC bytes must equal
the oracle's and ours.  (Valid for Merkle-shaped inputs only: zero lookup table, zero q_m / custom selectors — the
reference's native code has masked deviations elsewhere)  Skipped when the library was not built."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libzprize_ref.so")


@pytest.mark.parametrize("height", [4, 5])
def test_reference_native_prover_matches(pkg, gpu_lib, oracle, tmp_path, height):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libzprize_ref.so not built (oracle/build_pnp_ref.sh needs /root/reference)")
    out = str(tmp_path / "ref_proof.npy")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_pnp_reference.py"), "--height", str(height), "--out", out],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    if height != 4 and r.returncode != 0:
        # HEIGHT=4 (BASELINE.json configs[0]) is the pinned case; at some other sizes the reference's native code
        # crashes on its own (observed: SIGSEGV at HEIGHT=6 on sm_100) — that is not a parity failure of ours.
        pytest.skip("reference native prover crashed at HEIGHT=%d (rc=%d)" % (height, r.returncode))
    assert r.returncode == 0 and os.path.exists(out), r.stdout[-2000:]
    ref_proof = np.load(out)
    oracle_proof = np.load(out.replace(".npy", "_oracle.npy"))
    names = pkg.COMMITMENT_NAMES
    diff = [names[i] for i in range(19) if not np.array_equal(ref_proof[12 * i:12 * i + 12], oracle_proof[12 * i:12 * i + 12])]
    diff += [pkg.EVALUATION_NAMES[i] for i in range(26)
             if not np.array_equal(ref_proof[228 + 4 * i:232 + 4 * i], oracle_proof[228 + 4 * i:232 + 4 * i])]
    assert not diff, "reference native prover differs from oracle in: %s" % diff
    # and our library on the same inputs
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, 0, with_pk=False)
    ctx = pkg.ProverContext(oc.log_n, gpu_lib)
    ctx.load_srs(oc.srs())
    ctx.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    assert np.array_equal(ctx.prove(circ).to_words(), ref_proof)
    ctx.close()
    oc.close()
