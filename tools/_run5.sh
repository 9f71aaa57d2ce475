mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_t2.log
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench3.err
python bench.py --workload motor --steps 20 > gpurun_out/r2_bench3_motor.json 2> gpurun_out/r2_bench3_motor.err; echo "motor rc=$?"; tail -3 gpurun_out/r2_bench3_motor.err
