mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r1f_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f_smoke.log 2>&1; tail -1 gpurun_out/r1f_smoke.log
python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err; tail -2 gpurun_out/r1f_bench_n1.err; cut -c1-200 gpurun_out/r1f_bench_n1.json
