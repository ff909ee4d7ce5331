"""Four-step NTT sharded over 2 and 4 ranks on CPU (gloo + emulation layer): every rank holds a contiguous block of the
input and must end with the same contiguous block of the oracle's full-size transform, for all four transform kinds."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, log_n, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from conftest import load_package
    import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    lib = pkg.load_library(emu_path)
    orc = oracle_lib.load()
    n = 1 << log_n
    m = n // world
    x = orc.random_fr(1, n)
    ctx = pkg.ProverContext(8, lib)

    def alltoall(send, recv, nbytes):
        # "device" memory is host memory under emulation; emulate all-to-all with an all-gather of the send buffers
        sbuf = (ctypes.c_uint8 * (nbytes * world)).from_address(send)
        t = torch.frombuffer(sbuf, dtype=torch.uint8).clone()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        rbuf = (ctypes.c_uint8 * (nbytes * world)).from_address(recv)
        r = torch.frombuffer(rbuf, dtype=torch.uint8)
        for q in range(world):
            r[q * nbytes:(q + 1) * nbytes] = parts[q][rank * nbytes:(rank + 1) * nbytes]

    ok = True
    for kind in range(4):
        got = ctx.ntt_sharded(kind, log_n, rank, world, x[rank * m:(rank + 1) * m].copy(), alltoall)
        want = orc.ntt(kind, x)[rank * m:(rank + 1) * m]
        ok = ok and np.array_equal(got, want)
    open(os.path.join(out_dir, "ok_%d" % rank), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 10), (4, 11)])
def test_four_step_ntt_sharded(pkg, oracle, tmp_path, world, log_n):
    emu_path = pkg._build.build_emu()
    mp.spawn(_worker, args=(world, _free_port(), emu_path, log_n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "ok_%d" % r)).read() == "1"
