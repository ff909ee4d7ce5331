mkdir -p gpurun_out
python tools/bench_msm.py --logs 22 --iters 3 --batch 1 > gpurun_out/r1f_k32.log 2>&1; python tools/bench_msm.py --logs 22 --iters 3 --batch 6 >> gpurun_out/r1f_k32.log 2>&1; cut -c1-420 gpurun_out/r1f_k32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r1f_bench_n1.json')); print(d['value'], d['e2e']['value'], d['proof'], d['phase_ms_per_step'])"
