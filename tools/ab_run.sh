#!/bin/bash
# A/B of library build variants on one GPU box: tools/ab_run.sh <tag> <variant> [<variant> ...]   ("base" = product build)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset SPB_LIB_PATH; else export SPB_LIB_PATH=$PWD/spectre_b200/libspectre_b200_$v.so; fi
  echo "=== variant $v" >> gpurun_out/${tag}_ab.log
  python tools/microbench.py modmul_quick >> gpurun_out/${tag}_ab.log 2>&1
  python bench.py --steps 40 --warmup 5 --no-prove --no-cpu-baseline >> gpurun_out/${tag}_ab.log 2>&1
done
