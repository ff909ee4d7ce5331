mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_msm_sharded.py --logs 16,22 --iters 3 --batch 1 > gpurun_out/r2q_dbg.log 2>&1; tail -25 gpurun_out/r2q_dbg.log
