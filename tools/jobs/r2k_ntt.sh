mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -q -m gpu -x > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
(echo "== register radix-8 pass kernel (default)"; python tools/bench_msm.py --logs 16,18,20,22,24,25 --no-msm --ntt 1 --iters 5; echo "== smem radix-2 pass kernel (ZP_NTT_REG=0)"; ZP_NTT_REG=0 python tools/bench_msm.py --logs 16,22,25 --no-msm --ntt 1 --iters 5) > gpurun_out/r2k_ntt_sweep.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2k_ntt_sweep.log"):
    if l.startswith("=="): print(l.strip())
    elif l.startswith("{"):
        d=json.loads(l); print(d["log_n"], d["kind"], round(d["ms"],4))
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-drop-in > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['phase_ms_per_step'], d['roofline_ntt']['ms'], d['proof']['equals_pinned_oracle_proof'])"
