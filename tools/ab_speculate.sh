#!/bin/bash
# A/B of the speculative first line-search trial in ttmpc_solve_kernel ("Speculative first trial").
# Build the experiment library HERE first (nvcc cross-compiles; built .so files travel with the gpurun snapshot):
#   TTMPC_NVCC_FLAGS="-DTTMPC_SPECULATION=1" TTMPC_BUILD_OUT=$PWD/tools/libttmpc_spec.so \
#     python -c "from car_trailer_mpc_b200.build import build_library; build_library(force=True)"
# (add -DTTMPC_SPEC_DEBUG to print every rejected speculative step), then on the GPU box:  bash tools/ab_speculate.sh
set -u
SPEC=$PWD/tools/libttmpc_spec.so
O=gpurun_out/ab_speculate.txt
mkdir -p gpurun_out; : > $O
for B in 65536 4096; do
  echo "== shipped library (classic trial sweeps), B=$B" >> $O
  timeout 60 python tools/quick_bench.py $B 40 5 2>&1 | tail -1 >> $O
  for mode in 0 1 2 3; do
    echo "== experiment library, TTMPC_SPECULATE=$mode, B=$B" >> $O
    TTMPC_LIB=$SPEC TTMPC_SPECULATE=$mode TTMPC_DEBUG_RESTARTS=1 timeout 60 python tools/quick_bench.py $B 40 5 2>&1 | grep -v "^fp64" | uniq | tail -3 >> $O
  done
done
TTMPC_LIB=$SPEC TTMPC_SPECULATE=1 timeout 90 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2 >> $O
cat $O
