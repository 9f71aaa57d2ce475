#!/bin/bash
# Build tuning variants of libartist_b200.so on the CPU box (nvcc cross-compiles): only trace.cu is recompiled,
# the other objects come from the regular build.  usage: tools/build_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags ...]
# SRC=nurbs tools/build_variants.sh ... recompiles nurbs.cu instead (default: trace).
# The variants land in artist_b200/lib/variants/<name>.so and are selected at run time with AB200_LIB=<path>.
set -e
SRC=${SRC:-trace}
cd "$(dirname "$0")/.."
python -c "from artist_b200 import _build; _build.build()"
mkdir -p artist_b200/lib/variants
pids=()
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    extra=""
    if [ $SRC = trace ]; then   # the trace kernels are two translation units (trace.cu = forward half, trace_bwd.cu)
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
        -c artist_b200/csrc/trace_bwd.cu -o artist_b200/lib/variants/trace_bwd_$name.o &
      extra=artist_b200/lib/variants/trace_bwd_$name.o
    fi
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
      -c artist_b200/csrc/$SRC.cu -o artist_b200/lib/variants/${SRC}_$name.o && wait &&
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o artist_b200/lib/variants/$name.so \
      artist_b200/lib/variants/${SRC}_$name.o $extra $(for o in trace trace_bwd nurbs kinematics blocking geometry flux sampling; do [ $o = $SRC ] || { [ $SRC = trace ] && [ $o = trace_bwd ]; } || echo artist_b200/lib/$o.o; done) &&
    rm -f artist_b200/lib/variants/${SRC}_$name.o $extra && echo "built $name ($flags)"
  ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
