mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r1c_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1c_smoke.log 2>&1; tail -2 gpurun_out/r1c_smoke.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1c_bench_reference_arm.json 2> gpurun_out/r1c_ref.err; cat gpurun_out/r1c_bench_reference_arm.json
python bench.py --steps 5 --warmup 3 > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err; cat gpurun_out/r1c_bench_n1.json
python tools/profile_proof.py > gpurun_out/r1c_prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches_height15.csv python tools/profile_proof.py > gpurun_out/r1c_ncu_launch.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ba_down0|ba_up0" -c 2 -o gpurun_out/r1c_prof_ba python tools/profile_proof.py > gpurun_out/r1c_ncu_ba.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ntt_pass" -s 15 -c 3 -o gpurun_out/r1c_prof_ntt python tools/profile_proof.py > gpurun_out/r1c_ncu_ntt.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"quotient_kernel|msm_rowcol|msm_accumulate" -c 3 -o gpurun_out/r1c_prof_misc python tools/profile_proof.py > gpurun_out/r1c_ncu_misc.log 2>&1
ls -la gpurun_out
