set -x
mkdir -p gpurun_out
python tools/profile_proof.py > gpurun_out/r1b_prof_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_height15.csv python tools/profile_proof.py > gpurun_out/r1b_ncu_launch.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ba_down0|ba_up0" -c 2 -o gpurun_out/r1b_prof_ba python tools/profile_proof.py > gpurun_out/r1b_ncu_ba.log 2>&1
ls -la gpurun_out/
