mkdir -p gpurun_out
(for t in 32 64 128; do echo "== rowcol threads $t"; ZP_MSM_ROWCOL_THREADS=$t python tools/bench_msm.py --logs 22 --iters 3 --batch 1; ZP_MSM_ROWCOL_THREADS=$t python tools/bench_msm.py --logs 22 --iters 3 --batch 6; done) > gpurun_out/r1d_rowcol.log 2>&1
cat gpurun_out/r1d_rowcol.log | cut -c1-400
python bench.py --steps 5 --warmup 3 > gpurun_out/r1d_bench_n1.json 2> gpurun_out/r1d_bench_n1.err; cat gpurun_out/r1d_bench_n1.json
