"""BASELINE config 3: G1 MSM sweep 2^16 .. 2^24 on 1 / 2 / 4 / 8 B200 (one process per GPU, torchrun):
  torchrun --nproc-per-node G tools/bench_msm_sharded.py --sizes 16,18,20,22,24 [--batch 4]
Every rank holds the SRS and the scalars; the MSM is split by bucket share (precomputed window tables, n >= 2^16) and the
192-byte partial sums are all-gathered and folded.  Time = CUDA events on every rank, MAX over ranks; the result is
checked against the known-trapdoor identity  MSM(s) = [sum_i s_i tau^i] G  on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_package  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,18,20,22")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    tau = orc.random_fr(7, 1)[0]
    for lg in [int(x) for x in args.sizes.split(",")]:
        n = 1 << lg
        x = orc.random_fr(2, n)
        ctx = pkg.ProverContext(max(lg, 6), lib)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.generate_srs(tau)
        if world > 1:
            bufs = {}

            def allgather(data):
                if len(data) not in bufs:
                    bufs[len(data)] = (torch.empty(len(data), dtype=torch.uint8, device="cuda"),
                                       torch.empty(len(data) * world, dtype=torch.uint8, device="cuda"))
                gi, go = bufs[len(data)]
                gi.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
                dist.all_gather_into_tensor(go, gi)
                return bytes(go.cpu().numpy())

            ctx.set_shard(rank, world, allgather)
        ctx.bench_alloc(0, n)
        ctx.bench_upload(0, x)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms, out = ctx.bench_commit_sharded(0, n, args.iters, args.batch)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            g = ctx.read_srs(1)[0]
            ok = bool(np.array_equal(out, orc.g1_mul(g, orc.poly_eval(x, tau))))
            print(json.dumps({"op": "msm_sharded", "log_n": lg, "gpus": world, "batch": args.batch, "ms": float(t[0]),
                              "ms_per_msm": float(t[0]) / args.batch, "points_per_s": n * args.batch / float(t[0]) * 1e3,
                              "trapdoor_identity_ok": ok, "mode": "points" if os.environ.get("ZP_SHARD_BUCKETS") == "0" else "buckets"}),
                  flush=True)
        ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
