mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi_rank.py tests/test_witness_synthesis.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2i_pytest.log 2>&1; tail -5 gpurun_out/r2i_pytest.log
(python tools/bench_msm.py --logs 22 --iters 3 --batch 4; python tools/bench_msm.py --logs 22 --iters 3 --batch 1; python tools/bench_msm.py --logs 20 --iters 3 --batch 1) > gpurun_out/r2i_msm_prefetch.log 2>&1; cat gpurun_out/r2i_msm_prefetch.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-drop-in > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2i_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['phase_ms_per_step'], d['roofline_msm_stage']['executed_frac'], d['proof']['equals_pinned_oracle_proof'])"
