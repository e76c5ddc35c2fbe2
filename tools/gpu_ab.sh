#!/usr/bin/env bash
# A/B of fused-kernel builds on one GPU box: per variant the wait/timeline trace at k = 3 and 8 and a short bench.
# usage: tools/gpu_ab.sh <outdir> <variant>...   (variant "main" = the in-tree library, else ab/lib_<variant>.so)
set -u
o=$1; shift
mkdir -p "$o"
for v in "$@"; do
  if [ "$v" = main ]; then unset RESNMTF_B200_LIB; else export RESNMTF_B200_LIB=$PWD/ab/lib_$v.so; fi
  for k in 3 8; do
    RESNMTF_FU_TIMELINE=1 python tools/profile_run.py --k $k --iters 30 > "$o/tl_${v}_k$k.txt" 2>&1
  done
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded --no-ksweep > "$o/ab_$v.json" 2> "$o/ab_$v.err"
done
unset RESNMTF_B200_LIB
