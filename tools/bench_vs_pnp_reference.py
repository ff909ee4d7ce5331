"""Same-box GPU column: the REFERENCE's own NTT / MSM operators (sppark-derived kernels behind
`Ntt/Intt/Ntt_coset/Intt_coset::forward` and `multi_scalar_mult`, lib/PLONK/utils/function.cu:249-290), compiled
unmodified for sm_100 (oracle/_ref/libzprize_ref.so + the extern "C" driver oracle/ref_ops.cu), timed on this B200 next
to our kernels on identical inputs, with the outputs compared byte for byte.

  python tools/bench_vs_pnp_reference.py --logs 16,18,20,22,24 > profiles/rXX_vs_pnp_reference.jsonl

Each size runs in a SUBPROCESS for the reference side: its native code exits / crashes the process on some inputs."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402

u64p = ctypes.POINTER(ctypes.c_uint64)


def p(a):
    return a.ctypes.data_as(u64p)


def ref_worker(op, lg, kind, iters, path):
    """Runs in the child: reference operator on the arrays stored in `path`; writes timing + output next to it."""
    ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_ops.so"))
    d = np.load(path)
    ms, med = ctypes.c_double(), ctypes.c_double()
    if op == "ntt":
        n = 1 << lg
        out = np.zeros((8 * n if kind == 2 else n, 4), dtype=np.uint64)
        ref.ref_ops_ntt.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        ref.ref_ops_ntt(kind, lg, p(d["x"]), p(out), iters, ctypes.byref(ms), ctypes.byref(med))
    else:
        out = np.zeros(18, dtype=np.uint64)
        ref.ref_ops_msm.argtypes = [ctypes.c_size_t, u64p, u64p, u64p, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        ref.ref_ops_msm(1 << lg, p(d["points"]), p(d["x"]), p(out), iters, ctypes.byref(ms), ctypes.byref(med))
    np.savez(path.replace(".npz", "_out.npz"), out=out, ms=ms.value, med=med.value)


def run_ref(op, lg, kind, iters, path):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", op, str(lg), str(kind), str(iters), path],
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900)
    outp = path.replace(".npz", "_out.npz")
    if r.returncode != 0 or not os.path.exists(outp):
        return None, None, "rc=%d %s" % (r.returncode, r.stderr[-300:])
    d = np.load(outp)
    os.remove(outp)
    return (float(d["ms"]), float(d["med"])), d["out"], None


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        ref_worker(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--logs", default="16,18,20,22")
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--tmp", default="/tmp")
    ap.add_argument("--no-msm", action="store_true", dest="no_msm")
    ap.add_argument("--no-ntt", action="store_true", dest="no_ntt")
    args = ap.parse_args()
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    logs = [int(x) for x in args.logs.split(",")]
    tau = orc.random_fr(7, 1)[0]
    names = ["ntt", "intt", "coset_ntt_8n", "coset_intt"]
    for lg in logs:
        n = 1 << lg
        x = orc.random_fr(1, n)
        path = os.path.join(args.tmp, "zp_ref_in_%d.npz" % lg)
        ctx = pkg.ProverContext(max(lg, 6), lib)  # > 2^23: operator-only context (SRS, MSM, NTT entry points)
        if not args.no_ntt:
            np.savez(path, x=x)
            big = 8 * n if lg + 3 <= 26 else n
            ctx.bench_alloc(0, big)
            ctx.bench_alloc(1, big)
            ctx.bench_upload(0, x)
            for kind in range(4):
                if kind == 2 and lg + 3 > 26:
                    continue
                # ours: kind 2 = coset NTT of n coefficients zero-padded to 8n (what the quotient round runs)
                our_lg = lg + 3 if kind == 2 else lg
                if kind == 2:
                    ours_ms = bench_coset_padded(ctx, lg, args.iters)
                    ours_out = None
                else:
                    ours_ms = ctx.bench_ntt(kind, lg, 0, 1, args.iters)
                    ours_out = ctx.bench_download(1, n)
                ref_ms, ref_out, err = run_ref("ntt", lg, kind, args.iters, path)
                same = None
                if ref_out is not None and ours_out is not None:
                    same = bool(np.array_equal(ref_out, ours_out))
                elif ref_out is not None and kind == 2:
                    xp = np.zeros((8 * n, 4), dtype=np.uint64)
                    xp[:n] = x
                    same = bool(np.array_equal(ref_out, ctx.ntt(2, xp))) if lg + 3 <= 23 else None
                print(json.dumps({"op": names[kind], "log_n": lg, "out_log_n": our_lg, "reference_ms_min": ref_ms and ref_ms[0],
                                  "reference_ms_median": ref_ms and ref_ms[1], "ours_ms": ours_ms,
                                  "ratio_vs_reference_min": (ref_ms[0] / ours_ms) if ref_ms else None, "same_bytes": same,
                                  "reference_error": err}), flush=True)
            os.remove(path)
        if not args.no_msm:
            ctx.generate_srs(tau)
            pts = ctx.read_srs(n)
            ctx.bench_alloc(2, n)
            ctx.bench_upload(2, x)
            ours_ms, ours_out, _ = ctx.bench_msm(2, n, args.iters, 1)
            np.savez(path, x=x, points=pts)
            ref_ms, ref_out, err = run_ref("msm", lg, 0, args.iters, path)
            os.remove(path)
            same = None
            if ref_out is not None:
                same = bool(np.array_equal(jac_to_affine(orc, ref_out), ours_out))
            print(json.dumps({"op": "msm", "log_n": lg, "reference_ms_min": ref_ms and ref_ms[0], "reference_ms_median": ref_ms and ref_ms[1],
                              "ours_ms": ours_ms, "ratio_vs_reference_min": (ref_ms[0] / ours_ms) if ref_ms else None,
                              "same_bytes": same, "reference_error": err}), flush=True)
        ctx.close()


def bench_coset_padded(ctx, lg, iters):
    """coset NTT of 2^lg coefficients onto the 8x larger coset (implicit zero padding) — device resident."""
    import ctypes as ct
    ms = ct.c_double()
    rc = ctx.lib.zp_bench_ntt_padded(ctx.h, 2, lg + 3, 1 << lg, 0, 1, iters, ct.byref(ms))
    if rc != 0:
        raise RuntimeError(ctx.lib.zp_last_error().decode())
    return ms.value


def jac_to_affine(orc, jac18):
    """(X, Y, Z) Jacobian Montgomery -> affine 12 words with the FFI encoding of infinity."""
    X, Y, Z = (jac18[6 * i:6 * i + 6].reshape(1, 6).copy() for i in range(3))
    if not Z.any():
        out = np.zeros(12, dtype=np.uint64)
        one = np.zeros(6, np.uint64)
        m, rr = np.zeros(6, np.uint64), np.zeros(6, np.uint64)
        inv = ctypes.c_uint64()
        orc.lib.zpo_fq_constants(p(m), p(one), p(rr), ctypes.cast(ctypes.byref(inv), u64p))
        out[6:] = one
        return out
    zi = orc.fq_op(3, Z)
    zi2 = orc.fq_op(2, zi, zi)
    zi3 = orc.fq_op(2, zi2, zi)
    return np.concatenate([orc.fq_op(2, X, zi2)[0], orc.fq_op(2, Y, zi3)[0]])


if __name__ == "__main__":
    main()
