#!/bin/bash
for kb in 100 64 48 40 32; do
  export AB200_NVCC_EXTRA="-DAB200_NURBS_BWD_KB=$kb"
  python -m artist_b200._build > /dev/null 2>&1 || { echo "build failed $kb"; continue; }
  python bench.py --steps 5 --warmup 3 --skip-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('kb $kb', 'step', round(d['ms_per_step'],3), 'nurbs_bwd', d['kernel_ms']['ab200_nurbs_bwd'], 'nurbs_fwd', d['kernel_ms']['ab200_nurbs_fwd'])"
done
