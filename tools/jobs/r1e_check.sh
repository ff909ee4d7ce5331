mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r1e_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1e_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r1e_bench_n1.json 2> gpurun_out/r1e_bench_n1.err; tail -3 gpurun_out/r1e_bench_n1.err; cat gpurun_out/r1e_bench_n1.json
