mkdir -p gpurun_out
for r in 4 5; do ZP_MSM_BA_ROUNDS=$r python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rounds', $r, d['value'], d['phase_ms_per_step'], d['proof']['sha256'][:12])"; done > gpurun_out/r1f_rounds.log 2>&1
cat gpurun_out/r1f_rounds.log
