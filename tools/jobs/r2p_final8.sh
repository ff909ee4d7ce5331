# round 2 final multi-GPU pass on one 8 x B200 box: NCCL multi-rank parity test, bench at 8 / 4 / 2, sharded MSM sweep
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi_rank.py -q -m gpu > gpurun_out/r2p_pytest_multi_rank_nccl.log 2>&1; tail -3 gpurun_out/r2p_pytest_multi_rank_nccl.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu-baseline; }
for g in 8 4 2; do run $g > gpurun_out/r2p_scale_n$g.json 2> gpurun_out/r2p_scale_n$g.err; tail -1 gpurun_out/r2p_scale_n$g.err; done
sweep() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 tools/bench_msm_sharded.py --logs 16,18,20,22,24 --iters 5 --batch $2 2>/dev/null | grep '^{'; }
(for g in 8 4 2; do sweep $g 1; done; python tools/bench_msm_sharded.py --logs 16,18,20,22,24 --iters 5 --batch 1 | grep '^{'; sweep 8 4; ZP_SHARD_BUCKETS=0 sweep 8 1 | sed 's/$/  # point-range mode/') > gpurun_out/r2p_msm_sharded_sweep.jsonl 2>&1
cat gpurun_out/r2p_msm_sharded_sweep.jsonl | cut -c1-200
python - <<'PY'
import json
for g in [8,4,2]:
    try:
        d=json.loads(open("gpurun_out/r2p_scale_n%d.json"%g).read().strip().splitlines()[-1])
        print(g, d["value"], d["e2e"]["value"], d["phase_ms_per_step"], d["proof"]["equals_pinned_oracle_proof"], d["proof"]["identical_on_all_ranks"])
    except Exception as e: print(g, "failed", e)
PY
