#!/bin/bash
# Build tuning variants of libartist_b200.so on the CPU box (nvcc cross-compiles): only trace.cu is recompiled,
# the other objects come from the regular build.  usage: tools/build_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags ...]
# The variants land in artist_b200/lib/variants/<name>.so and are selected at run time with AB200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python -c "from artist_b200 import _build; _build.build()"
mkdir -p artist_b200/lib/variants
pids=()
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
      -c artist_b200/csrc/trace.cu -o artist_b200/lib/variants/trace_$name.o &&
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o artist_b200/lib/variants/$name.so \
      artist_b200/lib/variants/trace_$name.o artist_b200/lib/nurbs.o artist_b200/lib/kinematics.o artist_b200/lib/blocking.o &&
    rm artist_b200/lib/variants/trace_$name.o && echo "built $name ($flags)"
  ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
