#!/bin/bash
# Round-end measurement batch on ONE B200 (run under gpurun): tests, bench lines, ncu launch list of the bench command,
# one `ncu --set full` capture of every kernel of a step.  Each ncu pass only after its command exited 0 without ncu.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err || exit 1
python bench.py --workload motor --steps 20 > gpurun_out/r2_bench_final_motor.json 2> gpurun_out/r2_bench_final_motor.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_final_reference.json 2>/dev/null
python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
python tools/profile_trace.py 2048 2 > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'trace_|nurbs_|bitmaps_per' --launch-skip 5 --launch-count 5 \
    -o gpurun_out/r2_final_all -f python tools/profile_trace.py 2048 2 > gpurun_out/r2_ncu_full.log 2>&1
grep "Profiling" gpurun_out/r2_ncu_full.log | cut -c1-80
