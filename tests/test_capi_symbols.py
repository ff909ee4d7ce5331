"""The C-ABI library loads without a GPU and exports every symbol include/zprize_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "zprize_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n.startswith("zp_") or n == "gen_proof"))


def test_header_symbols_exported(pkg):
    lib_path = pkg._build.build()
    lib = ctypes.CDLL(lib_path)
    declared = _declared_symbols()
    assert "gen_proof" in declared and len(declared) >= 25
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(pkg.EXPORTED_SYMBOLS) == declared


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product refuses to create a context (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    lib = pkg.load_library(pkg._build.build())
    assert lib.zp_device_available() == 0
    try:
        pkg.ProverContext(11, lib)
    except pkg.ZprizeError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("context creation must fail without a GPU")


def test_struct_layout_matches_ffi(pkg):
    # sizes implied by "Prize 1B/plonk-core/src/lib.rs":53-235
    assert ctypes.sizeof(pkg.CommitmentC) == 96
    assert ctypes.sizeof(pkg.ProofEvaluationsC) == 26 * 32
    assert ctypes.sizeof(pkg.ProofC) == 19 * 96 + 26 * 32 == 2656
    assert ctypes.sizeof(pkg.CircuitC) == 9 * 8
    assert ctypes.sizeof(pkg.ProverKeyC) == 44 * 8
    assert ctypes.sizeof(pkg.CommitKeyC) == 16
