#!/bin/bash
# Run on the GPU box: time the bench step with each pre-built library variant (tools/build_variants.sh).
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset AB200_LIB; else export AB200_LIB=artist_b200/lib/variants/$lib.so; fi
  python bench.py --steps 8 --warmup 3 --skip-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'step', round(d['ms_per_step'],3), 'fwd', d['kernel_ms']['ab200_trace_fwd'], 'bwd', d['kernel_ms']['ab200_trace_bwd'], 'nurbs', d['kernel_ms']['ab200_nurbs_fwd'], d['kernel_ms']['ab200_nurbs_bwd'])"
done
