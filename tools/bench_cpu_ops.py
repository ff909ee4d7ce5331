"""CPU column of the operator sweeps ("for NTT/MSM sweeps time the same CPU routines at the same sizes"):
the oracle's NTT family (radix-2, OpenMP, blst field products when oracle/_ref/libref_blst.so is present) and its MSM
(`blst_p1s_mult_pippenger` over thread ranges) on this host's cores, same inputs as tools/bench_msm.py and the NTT sweep
(uniform Fr seed 1 / seed 2, SRS tau seed 7).  One JSON line per (operator, size).  TEST INFRASTRUCTURE: executes oracle/."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log", type=int, default=16)
    ap.add_argument("--max-log", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    orc = oracle_lib.load()
    threads = len(os.sched_getaffinity(0))
    orc.lib.zpo_set_num_threads(threads)
    cpu = [ln.split(":")[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][:1]
    meta = {"threads": int(orc.lib.zpo_num_threads()), "blst": bool(orc.lib.zpo_blst_active()), "cpu": cpu[0] if cpu else "?"}
    srs, _ = orc.srs(7, 1 << args.max_log)
    for lg in range(args.min_log, args.max_log + 1, 2):
        n = 1 << lg
        x = orc.random_fr(1, n)
        for kind, name in ((0, "ntt"), (1, "intt"), (2, "coset_ntt"), (3, "coset_intt")):
            ts = []
            for _ in range(args.reps):
                buf = x.copy()
                t = time.perf_counter()
                orc.lib.zpo_ntt(kind, lg, oracle_lib._p(buf))
                ts.append(time.perf_counter() - t)
            print(json.dumps({"op": name, "log_n": lg, "ms_min": min(ts) * 1e3, "ms_median": float(np.median(ts)) * 1e3,
                              "elems_per_s": n / min(ts), **meta}), flush=True)
        s = orc.random_fr(2, n)
        ts = []
        for _ in range(args.reps):
            t = time.perf_counter()
            orc.msm(srs[:n], s)
            ts.append(time.perf_counter() - t)
        print(json.dumps({"op": "msm", "log_n": lg, "ms_min": min(ts) * 1e3, "ms_median": float(np.median(ts)) * 1e3,
                          "points_per_s": n / min(ts), **meta}), flush=True)


if __name__ == "__main__":
    main()
