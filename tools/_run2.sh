mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench2.err
python bench.py --workload motor --steps 20 > gpurun_out/r2_bench2_motor.json 2> gpurun_out/r2_bench2_motor.err; echo "motor rc=$?"; tail -5 gpurun_out/r2_bench2_motor.err
