mkdir -p gpurun_out
(echo "== whole MSM, batch 4"; python tools/bench_msm.py --logs 22 --iters 3 --batch 4
 echo "== one rank's bucket slice of 8, batch 4"; ZP_BENCH_BUCKET_WORLD=8 python tools/bench_msm.py --logs 22 --iters 3 --batch 4
 echo "== one rank's bucket slice of 8, batch 6"; ZP_BENCH_BUCKET_WORLD=8 python tools/bench_msm.py --logs 22 --iters 3 --batch 6
 echo "== one rank's bucket slice of 2, batch 4"; ZP_BENCH_BUCKET_WORLD=2 python tools/bench_msm.py --logs 22 --iters 3 --batch 4
 echo "== point slice 2^19, batch 4"; python tools/bench_msm.py --logs 19 --iters 3 --batch 4
 echo "== point slice 2^19, batch 6"; python tools/bench_msm.py --logs 19 --iters 3 --batch 6) > gpurun_out/r2g_msm_slice.log 2>&1
cat gpurun_out/r2g_msm_slice.log
