"""One resident HEIGHT=h proof inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    oc = oracle_lib.OracleCircuit(orc, args.height, 42, 7, 0, with_pk=False, with_srs=False)
    ctx = pkg.ProverContext(oc.log_n, lib)
    ctx.generate_srs(oc.tau())
    ctx.preprocess(oc.selector_evals(), oc.tables())
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    for _ in range(args.warmup):
        ctx.prove(circ)
    ctx.upload_witness(circ)
    lib.zp_profiler_range(1)
    ctx.prove_resident()
    lib.zp_profiler_range(0)
    print("profiled proof done; timing", ctx.last_timing())


if __name__ == "__main__":
    main()
