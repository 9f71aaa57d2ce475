mkdir -p gpurun_out
export AB200_BENCH_MOTOR_TARGET=0
python tools/profile_trace.py 2048 2 motor > /dev/null 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:trace_bwd --launch-skip 1 --launch-count 1 -o gpurun_out/r2_motor_b -f python tools/profile_trace.py 2048 2 motor > gpurun_out/ncu_r2_motor_b.log 2>&1
tail -3 gpurun_out/ncu_r2_motor_b.log
