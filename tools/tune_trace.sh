#!/bin/bash
for cfg in "1024 1024 200 1" "512 512 100 1" "512 1024 100 1"; do
  set -- $cfg
  export AB200_WIN_KB=$3
  export AB200_NVCC_EXTRA="-DAB200_FWD_THREADS=$1 -DAB200_BWD_THREADS=$2 -DAB200_WIN_KB=$3 -DAB200_RAY_UNROLL=$4"
  python -m artist_b200._build > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  for bump in 0 1e-4 5e-4; do
  AB200_BENCH_BUMP=$bump python bench.py --steps 5 --warmup 3 --skip-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$cfg bump $bump', 'step', round(d['ms_per_step'],3), 'fwd', d['kernel_ms']['ab200_trace_fwd'], 'bwd', d['kernel_ms']['ab200_trace_bwd'])"
  AB200_BENCH_BUMP=$bump python tools/window_stats.py 2048 2>/dev/null | tail -1
  done
done
