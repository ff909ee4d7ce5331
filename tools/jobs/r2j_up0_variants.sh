mkdir -p gpurun_out
(for v in "" "ZP_BA_PF=1" "ZP_BA_PF=2" "ZP_BA_PF=4" "ZP_BA_NG=2" "ZP_BA_NG=4"; do echo "== ${v:-default}"; env $v python tools/bench_msm.py --logs 22 --iters 3 --batch 4; done) > gpurun_out/r2j_msm_up0_variants.log 2>&1
cat gpurun_out/r2j_msm_up0_variants.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' ')
    elif l.startswith('{'):
        d = json.loads(l); print(d['ms'], d['breakdown_ms']['batch_affine'])"
