mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(for v in "" "ZP_MSM_FUSED_SCATTER=1"; do echo "== ${v:-default} whole"; env $v python tools/bench_msm.py --logs 22 --iters 3 --batch 4; echo "== ${v:-default} share of 8"; env $v ZP_BENCH_BUCKET_WORLD=8 ZP_BENCH_BUCKET_RANK=3 python tools/bench_msm.py --logs 22 --iters 3 --batch 4; done) > gpurun_out/r2s_msm_fused_scatter.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2s_msm_fused_scatter.log"):
    if l.startswith("=="): print(l.strip(), end=" ")
    elif l.startswith("{"):
        d=json.loads(l); b=d["breakdown_ms"]; print(round(d["ms"],2), "digits", round(b["digits"],3), "scatter", round(b["scatter"],3))
PY
ZP_MSM_FUSED_SCATTER=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "msm or gen_proof_byte" 2>&1 | tail -2
