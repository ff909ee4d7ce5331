mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_ntt_sharded.py > gpurun_out/r1f_ntt_sharded_n8.log 2> gpurun_out/r1f_ntt_sharded_n8.err
tail -2 gpurun_out/r1f_ntt_sharded_n8.err; cut -c1-330 gpurun_out/r1f_ntt_sharded_n8.log
