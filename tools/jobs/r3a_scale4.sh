# round 2, final state on 4 GPUs (NCCL): the launch line the driver uses
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r3a_scale_n4.json 2> gpurun_out/r3a_scale_n4.err; echo "rc=$?"
tail -3 gpurun_out/r3a_scale_n4.err; cut -c1-200 gpurun_out/r3a_scale_n4.json
python - <<'PY'
import json
for ln in open("gpurun_out/r3a_scale_n4.json"):
    if ln.startswith("{"):
        d = json.loads(ln); print("value", d["value"], "e2e", d["e2e"]["value"], d["phase_ms_per_step"], d["proof"])
PY
