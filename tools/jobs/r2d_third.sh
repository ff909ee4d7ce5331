mkdir -p gpurun_out
python -m pytest tests -q -m gpu --durations=10 > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -16 gpurun_out/r2d_pytest_gpu.log
timeout 300 python tools/run_pnp_reference.py --height 5 --lib libzprize_ref_patched.so --out /tmp/refp_5.npy > gpurun_out/r2d_ref_patched_h5.log 2>&1
grep -v "^$" gpurun_out/r2d_ref_patched_h5.log | head -40 | cut -c1-220
