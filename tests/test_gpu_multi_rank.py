"""Multi-rank gen_proof ON REAL DEVICES, witnessed by pytest (not only by bench.py): torchrun spawns G processes that prove
one circuit cooperatively — commitments sharded by MSM bucket share, witness / quotient round exchanged on the device — and
the proof must equal the CPU oracle's bytes.  With >= G GPUs the collectives are NCCL; on a one-GPU box the ranks share
the device and exchange through gloo, which exercises exactly the same library paths."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# HEIGHT=9 (N = 2^16) takes the production MSM route: precomputed window tables, bucket shares, batch-affine rounds
@pytest.mark.parametrize("world,height,lookups,kind", [(2, 9, 0, 0), (2, 5, 16, 0), (4, 0, 12, 1), (8, 7, 0, 0)])
def test_sharded_proof_on_devices_equals_oracle(gpu_lib, oracle, tmp_path, world, height, lookups, kind):
    out = str(tmp_path / "proof.npy")
    env = dict(os.environ, ZP_DEAL_MIN_LOG="0", ZP_SHARD_BUCKETS_MIN_LOG="0")  # bucket shares / dealing also at these small sizes
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "run_sharded_proof.py"), "--height", str(height),
           "--lookups", str(lookups), "--kind", str(kind), "--out", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert r.returncode == 0 and os.path.exists(out), r.stdout[-3000:]
    oc = oracle_lib.OracleCircuit(oracle, height, 42, 7, lookups, kind=kind)
    ref, _ = oc.prove()
    oc.close()
    assert np.array_equal(np.load(out), ref)
