mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r1b_pytest_ntt.log 2>&1; tail -2 gpurun_out/r1b_pytest_ntt.log
(echo "== direct twiddle tables (default)"; python tools/bench_msm.py --logs 16,18,20,22,24,25 --no-msm --ntt 1; echo "== two-level lookup only (ZP_NTT_TW_MAX_LOG=0)"; ZP_NTT_TW_MAX_LOG=0 python tools/bench_msm.py --logs 22,25 --no-msm --ntt 1) > gpurun_out/r1b_ntt_sweep.log 2>&1
cat gpurun_out/r1b_ntt_sweep.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms_per_step'], d['roofline_ntt'])"
