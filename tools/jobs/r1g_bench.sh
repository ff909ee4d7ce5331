mkdir -p gpurun_out
python bench.py > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench_n1.err; tail -2 gpurun_out/r1g_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r1g_bench_n1.json')); print(d['value'], d['e2e']['value'], d['per_step'], d['proof']['verifier_accepts'])"
