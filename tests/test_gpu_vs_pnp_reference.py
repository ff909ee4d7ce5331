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
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
UNMODIFIED, PATCHED = "libzprize_ref.so", "libzprize_ref_patched.so"


# HEIGHT=4 (BASELINE.json configs[0]) runs the UNMODIFIED reference.  Above it the unmodified library dies on this box from
# its own double destruction of shared buffers (oracle/build_pnp_ref.sh, profiles/r02u_pnp_reference_fake_cudart.log), so the
# larger sizes — up to N = 2^17, where our prover is on its production MSM / NTT routes — run the library with that one
# defect patched at build time.
@pytest.mark.parametrize("height,lib", [(4, UNMODIFIED), (5, UNMODIFIED), (5, PATCHED), (6, PATCHED), (8, PATCHED), (10, PATCHED)])
def test_reference_native_prover_matches(pkg, gpu_lib, oracle, tmp_path, height, lib):
    if not os.path.exists(os.path.join(REF_DIR, lib)):
        pytest.skip("oracle/_ref/%s not built (oracle/build_pnp_ref.sh needs /root/reference)" % lib)
    out = str(tmp_path / "ref_proof.npy")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_pnp_reference.py"), "--height", str(height), "--out", out,
                        "--lib", lib], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    if lib == UNMODIFIED and height != 4 and r.returncode != 0:
        pytest.skip("unmodified reference native prover crashed at HEIGHT=%d (rc=%d): its own defect, see above" % (height, r.returncode))
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
