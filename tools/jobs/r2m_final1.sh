# round 2 final single-GPU pass: tests, smoke, both bench arms, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2m_pytest_gpu.log 2>&1; tail -14 gpurun_out/r2m_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -2 gpurun_out/r2m_bench_n1.err; cut -c1-1500 gpurun_out/r2m_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err; cut -c1-400 gpurun_out/r2m_bench_reference.json
python tools/profile_proof.py > gpurun_out/r2m_prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches_height15.csv python tools/profile_proof.py > gpurun_out/r2m_ncu_launch.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ba_down0" -c 20 -o gpurun_out/r2m_prof_down0 python tools/profile_proof.py > gpurun_out/r2m_ncu_down0.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ba_up0" -c 4 -o gpurun_out/r2m_prof_up0 python tools/profile_proof.py > gpurun_out/r2m_ncu_up0.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ntt_pass|quotient_kernel" -s 15 -c 4 -o gpurun_out/r2m_prof_ntt_quot python tools/profile_proof.py > gpurun_out/r2m_ncu_ntt.log 2>&1
ls -la gpurun_out | tail -12
