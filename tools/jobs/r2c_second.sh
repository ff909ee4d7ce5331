# round 2, second GPU pass: device sigma construction + pairing verifier on the box, patched reference at HEIGHT=5/6/15,
# same-box operator sweep (warm tables)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --durations=10 > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -18 gpurun_out/r2c_pytest_gpu.log
for h in 5 6; do
  echo "== patched reference HEIGHT=$h"; timeout 300 python tools/run_pnp_reference.py --height $h --lib libzprize_ref_patched.so --out /tmp/refp_$h.npy 2>&1 | grep -v "^$" | tail -6
done > gpurun_out/r2c_ref_patched.log 2>&1
(echo "== patched reference HEIGHT=15 (3 calls)"; timeout 1500 python tools/run_pnp_reference.py --height 15 --repeat 3 --lib libzprize_ref_patched.so --out /tmp/refp_15.npy 2>&1 | grep -v "^$" | tail -12) >> gpurun_out/r2c_ref_patched.log 2>&1
cat gpurun_out/r2c_ref_patched.log
timeout 900 python tools/bench_vs_pnp_reference.py --logs 16,18,20,22,24 --iters 3 > gpurun_out/r2c_vs_pnp_reference.jsonl 2> gpurun_out/r2c_vs_pnp_reference.err; cat gpurun_out/r2c_vs_pnp_reference.jsonl; tail -3 gpurun_out/r2c_vs_pnp_reference.err
