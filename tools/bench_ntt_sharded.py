"""Four-step sharded NTT on G GPUs (torchrun, NCCL all-to-all over NVLink): parity against the single-GPU transform,
then device-resident timing.  Launch: python -m torch.distributed.run --nproc-per-node G tools/bench_ntt_sharded.py"""
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


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logs", default="20,22,25")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    ctx = pkg.ProverContext(10, lib)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def alltoall(send, recv, nbytes):
        s = torch.as_tensor(_DevMem(send, nbytes * world), device="cuda")
        r = torch.as_tensor(_DevMem(recv, nbytes * world), device="cuda")
        dist.all_to_all_single(r, s)

    for lg in [int(v) for v in args.logs.split(",")]:
        n = 1 << lg
        m = n // world
        x = orc.random_fr(1, n)
        ok = None
        if lg <= 22:  # parity against the single-GPU transform of the same library (itself checked against the oracle)
            ok = True
            for kind in range(4):
                got = ctx.ntt_sharded(kind, lg, rank, world, x[rank * m:(rank + 1) * m].copy(), alltoall)
                want = ctx.ntt(kind, x)[rank * m:(rank + 1) * m]
                ok = ok and bool(np.array_equal(got, want))
        for s in range(4):
            ctx.bench_alloc(s, m)
        ctx.bench_upload(0, x[rank * m:(rank + 1) * m].copy())
        res = {}
        for kind, name in [(0, "ntt"), (2, "coset_ntt"), (3, "coset_intt")]:
            ctx.bench_ntt_sharded(kind, lg, rank, world, [0, 1, 2, 3], 2, alltoall)
            dist.barrier()
            ms = ctx.bench_ntt_sharded(kind, lg, rank, world, [0, 1, 2, 3], args.iters, alltoall)
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = float(t[0])
        if rank == 0:
            print(json.dumps({"op": "ntt_sharded_four_step", "n_gpus": world, "log_n": lg, "parity_vs_single_gpu": ok, "ms": res,
                              "elems_per_s": {k: n / v * 1e3 for k, v in res.items()}}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
