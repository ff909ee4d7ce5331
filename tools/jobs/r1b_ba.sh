set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch_affine or msm" > gpurun_out/r1b_pytest_ba.log 2>&1
for r in 0 1 2 3 4 5 6; do ZP_MSM_BA_ROUNDS=$r python tools/bench_msm.py --logs 22 --iters 3; done > gpurun_out/r1b_msm_ba.log 2>&1
python tools/bench_msm.py --logs 16,18,20,22 --iters 3 --dist witness > gpurun_out/r1b_msm_witness.log 2>&1
ZP_MSM_BA_ROUNDS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1b_ba_launches.csv python tools/bench_msm.py --logs 22 --iters 1 > gpurun_out/r1b_ba_ncu.log 2>&1
tail -3 gpurun_out/r1b_pytest_ba.log; cat gpurun_out/r1b_msm_ba.log
