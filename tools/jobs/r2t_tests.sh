mkdir -p gpurun_out
python -m pytest tests/test_custom_gates.py tests/test_gpu_fullsize_parity.py -q -m gpu --durations=5 > gpurun_out/r2t_pytest.log 2>&1; tail -10 gpurun_out/r2t_pytest.log
