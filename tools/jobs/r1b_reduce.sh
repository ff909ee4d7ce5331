mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r1b_pytest_red.log 2>&1; tail -2 gpurun_out/r1b_pytest_red.log
(python tools/bench_msm.py --logs 16,18,20,22 --iters 3 --batch 1; python tools/bench_msm.py --logs 22 --iters 3 --batch 4; python tools/bench_msm.py --logs 22 --iters 3 --batch 6) > gpurun_out/r1b_msm_red.log 2>&1
cat gpurun_out/r1b_msm_red.log
python tools/int_pipe.py > gpurun_out/r1b_int_pipe.log 2>&1; cat gpurun_out/r1b_int_pipe.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms_per_step'], d['roofline']['avg_ms_per_commitment'])"
