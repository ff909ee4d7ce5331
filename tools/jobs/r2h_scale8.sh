mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 5 --warmup 3 --no-cpu-baseline; }
run 8 > gpurun_out/r2h_n8.json 2> gpurun_out/r2h_n8.err; tail -2 gpurun_out/r2h_n8.err
run 4 > gpurun_out/r2h_n4.json 2> gpurun_out/r2h_n4.err; tail -2 gpurun_out/r2h_n4.err
run 2 > gpurun_out/r2h_n2.json 2> gpurun_out/r2h_n2.err; tail -2 gpurun_out/r2h_n2.err
python - <<'PY'
import json
for f in ["n8","n4","n2"]:
    try:
        d=json.loads(open("gpurun_out/r2h_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["phase_ms_per_step"], d["proof"]["equals_pinned_oracle_proof"], d["proof"]["identical_on_all_ranks"])
    except Exception as e: print(f, "failed", e)
PY
