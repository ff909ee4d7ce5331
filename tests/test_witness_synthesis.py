"""Device witness synthesis for the Poseidon-Merkle circuit (zp_prover_synthesize_merkle_witness, csrc/witness.cu) against
the oracle's composer: the same leaves / hash parameters / blinding values must give byte-identical wire columns, and a
proof made from the synthesized witness must equal the oracle's proof.  Reference gadget: constraint_system/hash.rs:20-127,
plonk-hashing/src/poseidon/zprize_constraints.rs:141-265, merkle-tree/src/lib.rs:41-59."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import _p


def merkle_inputs(oracle, height, seed):
    import ctypes
    oracle.lib.zpo_merkle_inputs.argtypes = [ctypes.c_int, ctypes.c_uint64, oracle_lib.u64p, oracle_lib.u64p, oracle_lib.u64p]
    blind = np.zeros((8, 4), dtype=np.uint64)
    leaves = np.zeros((1 << (height - 1), 4), dtype=np.uint64)
    params = np.zeros((201, 4), dtype=np.uint64)
    oracle.lib.zpo_merkle_inputs(height, seed, _p(blind), _p(leaves), _p(params))
    return blind, leaves, params


def _check(pkg, lib, oracle, height, prove):
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, 0, with_pk=prove, with_srs=prove)
    blind, leaves, params = merkle_inputs(oracle, height, 42)
    ctx = pkg.ProverContext(oc.log_n, lib)
    if prove:
        ctx.load_srs(oc.srs())
        ctx.preprocess(oc.selector_evals(), oc.tables())
    root = ctx.synthesize_merkle_witness(height, leaves, params, blind)
    got = ctx.read_witness()
    want = oc.wires()
    assert got[0].shape[0] == oc.cs_n
    for k in range(4):
        assert np.array_equal(got[k], want[k]), "wire column %d differs" % k
    # public input = -root
    neg_root = oracle.fr_op(4, oracle.fr_op(6, root.reshape(1, 4)))[0]
    assert np.array_equal(neg_root, oc.pi_canonical())
    if prove:
        ref, _ = oc.prove()
        assert np.array_equal(ctx.prove_resident().to_words(), ref)
    ctx.close()
    oc.close()


@pytest.mark.parametrize("height,prove", [(2, False), (4, False), (3, True)])
def test_emulated_witness_synthesis(pkg, emu_lib, oracle, height, prove):
    _check(pkg, emu_lib, oracle, height, prove)


@pytest.mark.gpu
@pytest.mark.parametrize("height,prove", [(4, True), (9, True), (12, False)])
def test_gpu_witness_synthesis(pkg, gpu_lib, oracle, height, prove):
    _check(pkg, gpu_lib, oracle, height, prove)
