"""Circuits that switch on every TurboPLONK widget of the quotient / linearisation path: range, logic (xor + and),
fixed-base scalar multiplication, curve addition, q_m, plus the degenerate key shapes q_c == 0 and q_arith == 0.

Gadget semantics follow the reference's composer ("Prize 1B/plonk-core/src/constraint_system/range.rs":27-211,
logic.rs:36-326, ecc/scalar_mul/fixed_base.rs:52-163, ecc/curve_addition/variable_base_gate.rs:25-98) with the inputs
of its own gadget tests; the oracle proves them on the CPU, the verifier restatement accepts, and the product path
(emulated here, on the device under -m gpu) must return the same 2656 bytes."""
import numpy as np
import pytest

import oracle_lib

KINDS = [(1, 0), (1, 12), (2, 0), (3, 0)]


@pytest.mark.parametrize("kind,n_lookup", KINDS)
def test_oracle_gadget_circuits_are_satisfied_and_verify(oracle, kind, n_lookup):
    oc = oracle_lib.OracleCircuit(oracle, 0, 42, 7, n_lookup, kind=kind)
    assert oc.satisfied()
    sel = oc.selector_evals()
    nz = {oracle_lib.PK_NAMES[i] for i in range(19) if sel[i].any()}
    if kind == 1:
        assert {"q_m", "range_selector", "logic_selector", "fixed_group_add_selector", "variable_group_add_selector"} <= nz
    if kind == 2:
        assert "q_c" not in nz and "q_m" in nz
    if kind == 3:
        assert "q_arith" not in nz and "range_selector" in nz
    proof, _ = oc.prove()
    ok, detail = oc.verify(proof)
    assert ok and detail == 3
    bad = proof.copy()
    bad[12 * 19 + 4 * 23] ^= 1  # a_next_eval: only the custom widgets read it
    assert not oc.verify(bad)[0]
    oc.close()


def test_jubjub_generator(oracle):
    assert oracle.lib.zpo_te_generator_on_curve() == 1


def _product_matches_oracle(pkg, lib, oracle, kind, n_lookup, use_host_pk, reps=0):
    oc = oracle_lib.OracleCircuit(oracle, reps, 42, 7, n_lookup, kind=kind)
    ref_proof, _ = oc.prove()
    ctx = pkg.ProverContext(oc.log_n, lib)
    ctx.load_srs(oc.srs())
    keep = None
    if use_host_pk:
        names = pkg.PK_POLY_NAMES + pkg.PK_SIGMA_NAMES
        keep = (oc.pk_coeffs(), oc.pk_evals(), oc.tables(), oc.linear_evaluations(), oc.v_h_coset_8n())
        ctx.load_pk(pkg.make_prover_key(dict(zip(names, keep[0])), dict(zip(names, keep[1])), keep[2], keep[3], keep[4]))
    else:
        ctx.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = ctx.prove(circ).to_words()
    assert np.array_equal(proof, ref_proof)
    assert oc.verify(proof)[0]
    ctx.close()
    oc.close()


@pytest.mark.parametrize("kind,n_lookup,use_host_pk", [(1, 0, False), (1, 12, True), (2, 0, False), (3, 0, False)])
def test_emulated_product_gadget_circuits(pkg, emu_lib, oracle, kind, n_lookup, use_host_pk):
    _product_matches_oracle(pkg, emu_lib, oracle, kind, n_lookup, use_host_pk)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n_lookup,use_host_pk", [(1, 0, False), (1, 0, True), (1, 12, False), (1, 12, True),
                                                       (2, 0, False), (2, 0, True), (3, 0, False), (3, 0, True)])
def test_gpu_gadget_circuits(pkg, gpu_lib, oracle, kind, n_lookup, use_host_pk):
    _product_matches_oracle(pkg, gpu_lib, oracle, kind, n_lookup, use_host_pk)


@pytest.mark.gpu
@pytest.mark.parametrize("use_host_pk", [False, True])
def test_gpu_gadget_circuit_through_production_routes(pkg, gpu_lib, oracle, use_host_pk):
    """70 repetitions of the gadget block + 300 plookup rows: N = 2^16, so the quotient kernel with CUSTOM and LOOKUP on, the
    custom linearisation terms and all ten 8N coset NTTs run together with the precomputed-table MSM, the batch-affine
    rounds and the multi-pass NTTs.  The oracle needs ~5-10 s for this proof."""
    _product_matches_oracle(pkg, gpu_lib, oracle, 1, 300, use_host_pk, reps=70)
