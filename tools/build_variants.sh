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
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
      -c artist_b200/csrc/$SRC.cu -o artist_b200/lib/variants/${SRC}_$name.o &&
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o artist_b200/lib/variants/$name.so \
      artist_b200/lib/variants/${SRC}_$name.o $(for o in trace nurbs kinematics blocking geometry flux sampling; do [ $o = $SRC ] || echo artist_b200/lib/$o.o; done) &&
    rm artist_b200/lib/variants/${SRC}_$name.o && echo "built $name ($flags)"
  ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
