#!/bin/bash
# A/B bench of alternative builds of the CUDA library (AG_LIB_PATH), run on the GPU box:
#   tools/ab_bench.sh "" build_ab/x.so build_ab/y.so     ("" = the in-tree default build)
# Appends one bench.py JSON line per variant to gpurun_out/ab.log.
mkdir -p gpurun_out
for v in "$@"; do
  if [ -n "$v" ]; then export AG_LIB_PATH=$PWD/$v; else unset AG_LIB_PATH; fi
  echo "== variant ${v:-default}" >> gpurun_out/ab.log
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $AB_ARGS >> gpurun_out/ab.log 2>&1
done
