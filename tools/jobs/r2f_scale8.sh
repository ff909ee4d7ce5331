mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 5 --warmup 3 --no-cpu-baseline; }
run 8 > gpurun_out/r2f_n8.json 2> gpurun_out/r2f_n8.err; tail -2 gpurun_out/r2f_n8.err
ZP_SHARD_BUCKETS=0 ZP_DEV_ALLGATHER=0 ZP_COSET_COPIES=0 ZP_DEAL_MIN_LOG=30 run 8 > gpurun_out/r2f_n8_round1_mode.json 2> gpurun_out/r2f_n8_round1_mode.err
run 4 > gpurun_out/r2f_n4.json 2> gpurun_out/r2f_n4.err; tail -2 gpurun_out/r2f_n4.err
python - <<'PY'
import json
for f in ["n8","n8_round1_mode","n4"]:
    try:
        d=json.loads(open("gpurun_out/r2f_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["phase_ms_per_step"], d["proof"]["equals_pinned_oracle_proof"], d["proof"]["identical_on_all_ranks"])
    except Exception as e: print(f, "failed", e)
PY
