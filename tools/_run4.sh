mkdir -p gpurun_out
python -m pytest tests/test_gpu_blocking.py tests/test_gpu_golden.py tests/test_gpu_api_end_to_end.py -q -x 2>&1 | tail -15
python tools/diag_motor.py 2048 5 2>&1 | tail -12
