# round 2: coset NTTs of the wires and of z forked onto a low-priority second stream (overlap with the commitment MSMs):
# A/B on the headline configuration, then the whole GPU suite and the full bench with the overlap on (default)
mkdir -p gpurun_out
for v in "off:ZP_NTT_OVERLAP=0" "on:ZP_NTT_OVERLAP=1" "on_equal_priority:ZP_NTT_OVERLAP_PRIO=-1"; do
  name="${v%%:*}"; kv="${v#*:}"
  env "$kv" timeout 600 python bench.py --steps 6 --warmup 3 --no-drop-in --no-cpu-baseline > gpurun_out/r2w_bench_$name.json 2> gpurun_out/r2w_bench_$name.err
  echo "$name rc=$?"; python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r2w_bench_%s.json" % sys.argv[1]))
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "phases", d["phase_ms_per_step"], "proof ok", d["proof"]["equals_pinned_oracle_proof"])
except Exception as e:
    print("no line:", e)
PY
done
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; tail -1 gpurun_out/r2w_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2w_bench_n1.json
