# where does the reference's own native prover crash above HEIGHT=4?  (native backtrace via oracle/libsegv_trace.so)
mkdir -p gpurun_out
for h in 5 5 6 8; do
  echo "== HEIGHT=$h plain"; timeout 300 python tools/run_pnp_reference.py --height $h --out /tmp/ref_$h.npy 2>&1 | tail -25; echo "rc=$?"
done > gpurun_out/r2b_ref_crash.log 2>&1
(echo "== HEIGHT=5 one core"; timeout 300 taskset -c 0 python tools/run_pnp_reference.py --height 5 --out /tmp/ref_5b.npy 2>&1 | tail -25) >> gpurun_out/r2b_ref_crash.log 2>&1
(echo "== HEIGHT=5 compute-sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_pnp_reference.py --height 5 --out /tmp/ref_5c.npy 2>&1 | tail -60) >> gpurun_out/r2b_ref_crash.log 2>&1
cat gpurun_out/r2b_ref_crash.log
python tools/bench_vs_pnp_reference.py --logs 24 --iters 2 --no-ntt > gpurun_out/r2b_vs_pnp_msm24.jsonl 2>&1; cat gpurun_out/r2b_vs_pnp_msm24.jsonl
