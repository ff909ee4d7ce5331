mkdir -p gpurun_out
sweep() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 tools/bench_msm_sharded.py --sizes 16,18,20,22,24 --iters 5 --batch $2 2>gpurun_out/r2r_sweep.err | grep '^{'; }
(for g in 8 4 2; do sweep $g 1; done; sweep 8 4; ZP_SHARD_BUCKETS=0 sweep 8 1) > gpurun_out/r2r_msm_sharded_sweep.jsonl 2>&1
cut -c1-175 gpurun_out/r2r_msm_sharded_sweep.jsonl; tail -3 gpurun_out/r2r_sweep.err
