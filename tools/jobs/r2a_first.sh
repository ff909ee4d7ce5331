# round 2, first GPU pass: new parity tests, reference crash backtrace, same-box sppark column, both bench arms
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --durations=15 > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -25 gpurun_out/r2a_pytest_gpu.log
timeout 600 cuda-gdb -batch -ex run -ex bt -ex "info sharedlibrary" --args python tools/run_pnp_reference.py --height 5 --out /tmp/ref5.npy > gpurun_out/r2a_ref_h5_gdb.log 2>&1; tail -40 gpurun_out/r2a_ref_h5_gdb.log
timeout 900 python tools/bench_vs_pnp_reference.py --logs 16,18,20,22 --iters 3 > gpurun_out/r2a_vs_pnp_reference.jsonl 2> gpurun_out/r2a_vs_pnp_reference.err; cat gpurun_out/r2a_vs_pnp_reference.jsonl; tail -5 gpurun_out/r2a_vs_pnp_reference.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -3 gpurun_out/r2a_bench_n1.err; cat gpurun_out/r2a_bench_n1.json
nproc; free -g | head -2; grep -m1 "model name" /proc/cpuinfo
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err; tail -3 gpurun_out/r2a_bench_reference.err; cat gpurun_out/r2a_bench_reference.json
