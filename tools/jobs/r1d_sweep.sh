mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size or msm_batch" > gpurun_out/r1d_pytest_fullsize.log 2>&1; tail -3 gpurun_out/r1d_pytest_fullsize.log
python tools/bench_msm.py --logs 23,24 --iters 2 > gpurun_out/r1d_msm_sweep_large.log 2>&1; cut -c1-420 gpurun_out/r1d_msm_sweep_large.log
