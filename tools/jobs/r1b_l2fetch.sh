mkdir -p gpurun_out
for g in default 32 64 128; do
  if [ $g = default ]; then unset ZP_L2_FETCH; else export ZP_L2_FETCH=$g; fi
  echo "== L2 fetch $g"
  python tools/bench_msm.py --logs 22 --iters 3 --batch 4
  ZP_MSM_BA_ROUNDS=0 python tools/bench_msm.py --logs 22 --iters 3
done > gpurun_out/r1b_l2fetch.log 2>&1
cat gpurun_out/r1b_l2fetch.log
