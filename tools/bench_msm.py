"""MSM / NTT sweep on the GPU with per-kernel breakdown (device-resident inputs). Results are checked against the
known-trapdoor identity commit(p) = [p(tau)] G at the smallest size."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logs", default="22")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--ntt", type=int, default=0)
    ap.add_argument("--no-msm", action="store_true", dest="no_msm")
    ap.add_argument("--batch", type=int, default=1, help="scalar vectors per MSM pipeline (the prover batches independent commitments)")
    ap.add_argument("--dist", default="uniform", help="uniform | witness (50%% zeros, 25%% values < 2^16, 25%% uniform; SURVEY 8d config 3)")
    args = ap.parse_args()
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    logs = [int(x) for x in args.logs.split(",")]
    lmax = max(logs)
    tau = orc.random_fr(7, 1)[0]
    n = 1 << min(lmax, 24)
    x = orc.random_fr(2, n)
    if args.dist == "witness":
        rng = np.random.default_rng(3)
        kind = rng.integers(0, 4, n)
        small = np.zeros((n, 4), dtype=np.uint64)
        small[:, 0] = rng.integers(1, 1 << 16, n)
        small = orc.fr_op(5, small)  # canonical -> Montgomery
        x[kind < 2] = 0
        x[kind == 2] = small[kind == 2]
    for lg in ([] if args.no_msm else logs):
        # one context per size so that window size / precomputed tables are tuned for that size
        m = 1 << lg
        c = pkg.ProverContext(max(lg, 6), lib)
        c.generate_srs(tau)
        c.bench_alloc(0, m)
        c.bench_upload(0, x[:m].copy())
        ms, out, bd = c.bench_msm(0, m, args.iters, args.batch)
        ctx = c
        line = {"op": "msm", "log_n": lg, "batch": args.batch, "ms": ms, "ms_per_msm": ms / args.batch,
                "points_per_s": m * args.batch / ms * 1e3, "breakdown_ms": bd, "ba_rounds": os.environ.get("ZP_MSM_BA_ROUNDS", "default"),
                "precomp": os.environ.get("ZP_MSM_PRECOMP", "1"), "dist": args.dist}
        if lg <= 16:
            srs = ctx.read_srs(m)
            line["matches_oracle"] = bool(np.array_equal(out, orc.msm(srs, x[:m].copy())))
        print(json.dumps(line), flush=True)
    if args.ntt:
        ctx = pkg.ProverContext(min(lmax, 23), lib)
        ctx.bench_alloc(0, 8 * n)
        ctx.bench_alloc(1, 8 * n)
        ctx.bench_upload(0, x)
        for lg in logs + ([lmax + 3] if lmax + 3 <= 26 and min(lmax, 23) + 3 >= lmax + 3 else []):
            for kind in range(4):
                ms = ctx.bench_ntt(kind, lg, 0, 1, args.iters)
                print(json.dumps({"op": "ntt", "kind": kind, "log_n": lg, "ms": ms, "elems_per_s": (1 << lg) / ms * 1e3,
                                  "algorithmic_GBps": 64.0 * (1 << lg) / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
