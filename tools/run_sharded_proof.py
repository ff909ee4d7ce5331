"""One rank of a multi-process sharded gen_proof on real GPUs (launched by torchrun from tests/test_gpu_multi_rank.py or
by hand):  torchrun --nproc-per-node G tools/run_sharded_proof.py --height H --out proof.npy [--lookups K] [--kind 1]

With >= G visible GPUs every rank takes its own device and the collectives are NCCL (the production path of bench.py);
with fewer, the ranks share device 0 and the collectives go through gloo with host staging — same library code, same
hooks (commitment partial sums all-gathered, witness slices / per-coset quotient coefficients exchanged on the device).
Rank 0 writes the proof; every rank asserts that it holds the same bytes."""
import argparse
import hashlib
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
    ap.add_argument("--height", type=int, default=6)
    ap.add_argument("--lookups", type=int, default=0)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    nccl = torch.cuda.device_count() >= world
    torch.cuda.set_device(local if nccl else 0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    lib = pkg.load_library()
    orc = oracle_lib.load()
    oc = oracle_lib.OracleCircuit(orc, args.height, 42, 7, args.lookups, with_pk=False, kind=args.kind)
    ctx = pkg.ProverContext(oc.log_n, lib)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.load_srs(oc.srs())
    ctx.preprocess(oc.selector_evals(), oc.tables())

    def allgather(data):
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        if nccl:
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return b"".join(bytes(o.cpu().numpy()) for o in out)

    def dev_bcast(ptr, nbytes, root):
        t = torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")
        if nccl:
            dist.broadcast(t, src=root)
        else:  # ranks share one GPU: stage through the host
            h = t.cpu()
            dist.broadcast(h, src=root)
            t.copy_(h)

    def dev_allgather(ptr, nbytes):
        whole = torch.as_tensor(_DevMem(ptr, nbytes * world), device="cuda")
        mine = whole[rank * nbytes:(rank + 1) * nbytes]
        if nccl:
            dist.all_gather_into_tensor(whole, mine)
        else:
            parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(parts, mine.cpu())
            whole.copy_(torch.cat(parts))

    ctx.set_shard(rank, world, allgather)
    ctx.set_device_broadcast(dev_bcast)
    ctx.set_device_allgather(dev_allgather)
    circ = pkg.make_circuit(oc.cs_n, oc.lookup_len, oc.pi_pos, oc.q_lookup(), oc.pi_canonical(), *oc.wires())
    proof = ctx.prove(circ).to_words()
    again = ctx.prove(circ).to_words()
    assert np.array_equal(proof, again)
    digest = hashlib.sha256(proof.tobytes()).digest()
    t = torch.tensor(list(digest), dtype=torch.uint8)
    if nccl:
        t = t.cuda()
    all_d = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(all_d, t)
    assert all(bool(torch.equal(x, t)) for x in all_d), "ranks hold different proofs"
    if rank == 0:
        np.save(args.out, proof)
        print("sharded proof written: world=%d backend=%s launches=%d" % (world, "nccl" if nccl else "gloo(one GPU)", lib.zp_launch_count()),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
