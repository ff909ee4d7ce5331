"""Sigma polynomials built ON THE DEVICE from the circuit's wire map (zp_sigma_from_wiring_host /
zp_prover_preprocess_wiring; reference: Permutation::compute_sigma_permutations + compute_permutation_lagrange,
"Prize 1B/plonk-core/src/permutation/mod.rs":101-166) against the oracle's columns, and against the two wire maps
whose sigmas the reference's own tests spell out (permutation/mod.rs:970-1190)."""
import numpy as np
import pytest

import oracle_lib
from test_reference_vectors import SIGMA_CASES, _encode


def _reference_vectors(pkg, lib, oracle):
    ctx = pkg.ProverContext(6, lib)  # N = 64 >= 4 gates: the first four rows must match the size-4 expectations' wiring
    for case in SIGMA_CASES:
        v = np.array(case["vars"], dtype=np.uint32)  # [wire][gate]
        vars_, cells = [], []
        for g in range(4):
            for w in range(4):
                vars_.append(v[w][g])
                cells.append((g << 2) | w)
        sig = ctx.sigma_from_wiring(np.array(vars_, np.uint32), np.array(cells, np.uint32), case["n_vars"])
        # expected WireData, encoded over the size-64 domain: K_wire * omega_64^gate
        w64, winv, ninv = (np.zeros(4, np.uint64) for _ in range(3))
        oracle.lib.zpo_fr_root_of_unity(6, oracle_lib._p(w64), oracle_lib._p(winv), oracle_lib._p(ninv))
        k = np.zeros((4, 4), dtype=np.uint64)
        k[:, 0] = [1, 7, 13, 17]
        k = oracle.fr_op(5, k)
        pw = [oracle.fr_op(5, np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]]
        for _ in range(3):
            pw.append(oracle.fr_op(2, pw[-1].reshape(1, 4), w64.reshape(1, 4))[0])
        for wire in range(4):
            for g in range(4):
                code = case["sigma"][wire][g]
                want = oracle.fr_op(2, k[code >> 28].reshape(1, 4), pw[code & 0xfffffff].reshape(1, 4))[0]
                assert np.array_equal(sig[wire][g], want), (wire, g)
    ctx.close()


def _matches_oracle(pkg, lib, oracle, height, kind, n_lookup=0):
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, n_lookup, with_pk=False, with_srs=False, kind=kind)
    v, c, nv = oc.wiring()
    ctx = pkg.ProverContext(oc.log_n, lib)
    sig = ctx.sigma_from_wiring(v, c, nv)
    sel = oc.selector_evals()
    for k in range(4):
        assert np.array_equal(sig[k], sel[15 + k]), k
    ctx.close()
    oc.close()


def test_emulated_sigma_reference_vectors(pkg, emu_lib, oracle):
    _reference_vectors(pkg, emu_lib, oracle)


@pytest.mark.parametrize("height,kind", [(3, 0), (0, 1), (0, 3)])
def test_emulated_sigma_matches_oracle(pkg, emu_lib, oracle, height, kind):
    _matches_oracle(pkg, emu_lib, oracle, height, kind)


def _prove_via_wiring(pkg, lib, oracle, height, kind, n_lookup):
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, n_lookup, kind=kind)
    ref_proof, _ = oc.prove()
    v, c, nv = oc.wiring()
    ctx = pkg.ProverContext(oc.log_n, lib)
    ctx.load_srs(oc.srs())
    ctx.preprocess_wiring(oc.selector_evals()[:15], v, c, nv, oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    assert np.array_equal(ctx.prove(circ).to_words(), ref_proof)
    ctx.close()
    oc.close()


def test_emulated_proof_with_device_sigmas(pkg, emu_lib, oracle):
    _prove_via_wiring(pkg, emu_lib, oracle, 3, 0, 12)


@pytest.mark.gpu
def test_gpu_sigma_reference_vectors(pkg, gpu_lib, oracle):
    _reference_vectors(pkg, gpu_lib, oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("height,kind", [(4, 0), (8, 0), (12, 0), (0, 1), (0, 2), (0, 3)])
def test_gpu_sigma_matches_oracle(pkg, gpu_lib, oracle, height, kind):
    _matches_oracle(pkg, gpu_lib, oracle, height, kind)


@pytest.mark.gpu
@pytest.mark.parametrize("height,kind,n_lookup", [(6, 0, 0), (4, 0, 24), (0, 1, 12)])
def test_gpu_proof_with_device_sigmas(pkg, gpu_lib, oracle, height, kind, n_lookup):
    _prove_via_wiring(pkg, gpu_lib, oracle, height, kind, n_lookup)
