# round 2, final state on 2 GPUs (NCCL): the launch line the driver uses
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2z_scale_n2.json 2> gpurun_out/r2z_scale_n2.err; echo "rc=$?"
tail -3 gpurun_out/r2z_scale_n2.err; cut -c1-200 gpurun_out/r2z_scale_n2.json
python - <<'PY'
import json
for ln in open("gpurun_out/r2z_scale_n2.json"):
    if ln.startswith("{"):
        d = json.loads(ln); print("value", d["value"], "e2e", d["e2e"]["value"], d["phase_ms_per_step"], d["proof"])
PY
