set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r1b_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1b_pytest_gpu.log
for b in 1 2 4 6; do python tools/bench_msm.py --logs 22 --iters 3 --batch $b; done > gpurun_out/r1b_msm_batch.log 2>&1
cat gpurun_out/r1b_msm_batch.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err; tail -2 gpurun_out/r1b_bench_n1.err; cat gpurun_out/r1b_bench_n1.json
