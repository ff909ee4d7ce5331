mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r1b_scale_n8.json 2> gpurun_out/r1b_scale_n8.err
tail -3 gpurun_out/r1b_scale_n8.err; cat gpurun_out/r1b_scale_n8.json
