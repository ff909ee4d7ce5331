# round 2 final single-GPU pass (second attempt: ncu reports are summarised on the box, only text comes back)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2n_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2n_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -1 gpurun_out/r2n_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; tail -2 gpurun_out/r2n_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err
python tools/profile_proof.py > gpurun_out/r2n_prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_height15.csv python tools/profile_proof.py > gpurun_out/r2n_ncu_launch.log 2>&1
cap() { # name, kernel regex, extra ncu args
  ncu --profile-from-start off --set full --clock-control none -k regex:"$2" $3 -o /tmp/$1 python tools/profile_proof.py > gpurun_out/r2n_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2n_ncu_$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/r2n_ncu_$1.md 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap down0 "ba_down0" "-c 20"
cap up0 "ba_up0" "-c 20"
cap ntt_quot "ntt_pass|quotient_kernel" "-s 15 -c 4"
cap msm_misc "msm_accumulate|msm_rowcol|msm_scatter|msm_digits" "-c 8"
du -sh gpurun_out; ls gpurun_out | tail -20
