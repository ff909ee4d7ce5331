# round 2, last single-GPU pass: the reference's native prover with its double destruction patched (HEIGHT=5 first, HEIGHT=15
# timed at the end), the whole GPU test suite, smoke, bench
mkdir -p gpurun_out
timeout 300 python tools/run_pnp_reference.py --height 5 --out /tmp/ref5.npy --lib libzprize_ref_patched.so > gpurun_out/r2u_ref_patched_h5.log 2>&1; echo "ref patched h5 rc=$?"; tail -2 gpurun_out/r2u_ref_patched_h5.log
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2u_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -1 gpurun_out/r2u_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2u_bench_n1.json
timeout 900 python tools/run_pnp_reference.py --height 15 --repeat 3 --out /tmp/ref15.npy --lib libzprize_ref_patched.so > gpurun_out/r2u_ref_patched_h15.log 2>&1; echo "ref patched h15 rc=$?"; grep -v "^\[" gpurun_out/r2u_ref_patched_h15.log | tail -6
