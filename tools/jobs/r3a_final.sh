# round 2, final single-GPU state: batch-inverted MSM window table, 8 staging threads, NTT overlap off by default, top of the inversion tree on the host + 8-lane slot table, work-list fold, SoA partial sums (r3a)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r3a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3a_smoke.log 2>&1; tail -1 gpurun_out/r3a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3a_bench_n1.json 2> gpurun_out/r3a_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3a_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "per_step", d["per_step"], "drop_in", {k: d["e2e_drop_in"][k] for k in ("cold_s", "cloned_key_s")}, "proof", d["proof"]["equals_pinned_oracle_proof"])
PY
