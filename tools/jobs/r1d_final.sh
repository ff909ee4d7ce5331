# Final evidence job of the round (1 GPU): tests, smoke, both bench arms, ncu launch list and full captures.
# The .ncu-rep files are summarised on the box (tools/ncu_summary.py) and deleted: gpurun_out/ is capped at 64 MiB.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r1d_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1d_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1; tail -2 gpurun_out/r1d_smoke.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1d_bench_reference_arm.json 2> gpurun_out/r1d_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r1d_bench_n1.json 2> gpurun_out/r1d_bench_n1.err; cat gpurun_out/r1d_bench_n1.json
python tools/bench_msm.py --logs 16,18,20,22 --iters 3 > gpurun_out/r1d_msm_sweep.log 2>&1
python tools/bench_msm.py --logs 16,18,20,22 --iters 3 --dist witness >> gpurun_out/r1d_msm_sweep.log 2>&1
python tools/bench_msm.py --logs 22 --iters 3 --batch 4 >> gpurun_out/r1d_msm_sweep.log 2>&1
python tools/bench_msm.py --logs 22 --iters 3 --batch 6 >> gpurun_out/r1d_msm_sweep.log 2>&1
python tools/profile_proof.py > gpurun_out/r1d_prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1d_launches_height15.csv python tools/profile_proof.py > gpurun_out/r1d_ncu_launch.log 2>&1
cap() {  # name, kernel regex, skip, count
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 python tools/profile_proof.py > gpurun_out/$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.md 2>> gpurun_out/$1.log
  rm -f gpurun_out/$1.ncu-rep
}
cap r1d_ncu_ba "ba_down0|ba_up0" 0 2
cap r1d_ncu_ntt "ntt_pass" 15 3
cap r1d_ncu_misc "quotient_kernel|msm_rowcol|msm_accumulate" 0 3
ls -la gpurun_out
