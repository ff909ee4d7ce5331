mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r1b_pytest_pad.log 2>&1; tail -2 gpurun_out/r1b_pytest_pad.log
(python tools/bench_msm.py --logs 22 --iters 3 --batch 1; python tools/bench_msm.py --logs 22 --iters 3 --batch 4; ZP_MSM_BA_ROUNDS=0 python tools/bench_msm.py --logs 22 --iters 3; ZP_MSM_BA_ROUNDS=4 python tools/bench_msm.py --logs 22 --iters 3 --batch 4) > gpurun_out/r1b_msm_pad.log 2>&1
cat gpurun_out/r1b_msm_pad.log
