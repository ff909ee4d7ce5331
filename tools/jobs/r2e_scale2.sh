# 2 GPUs: bucket-range MSM sharding vs point-range, all-gather hook; patched reference pointer diagnostics
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline; }
ZP_SHARD_BUCKETS=0 ZP_DEV_ALLGATHER=0 run > gpurun_out/r2e_n2_points.json 2> gpurun_out/r2e_n2_points.err; tail -2 gpurun_out/r2e_n2_points.err
run > gpurun_out/r2e_n2_buckets.json 2> gpurun_out/r2e_n2_buckets.err; tail -2 gpurun_out/r2e_n2_buckets.err
python - <<'PY'
import json
for f in ["points","buckets"]:
    try:
        d=json.loads(open("gpurun_out/r2e_n2_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["phase_ms_per_step"], d["proof"])
    except Exception as e: print(f, "failed", e)
PY
timeout 300 python tools/run_pnp_reference.py --height 5 --lib libzprize_ref_patched.so --out /tmp/refp_5.npy > gpurun_out/r2e_ref_patched_h5.log 2>&1
grep "ref-patch\|segv_trace\|equals" gpurun_out/r2e_ref_patched_h5.log | head -20
