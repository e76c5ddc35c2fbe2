#!/usr/bin/env bash
# A/B of one-pass kernel builds on one box.
# usage: tools/gpu_f2.sh <outdir> [variant ...]   variant: k1 (in-tree library, first generation forced) | k2 (in-tree,
#        second generation forced) | auto (in-tree, the plan's choice) | <name> (= ab/lib_<name>.so, second generation
#        forced) | <name>@1 (the same library, first generation forced)
set -u
o=${1:-gpurun_out/f2}; shift
mkdir -p "$o"
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu > "$o/pytest_fused.log" 2>&1
tail -3 "$o/pytest_fused.log"
for v in "$@"; do
  unset RESNMTF_B200_LIB RESNMTF_FUSED_KIND
  case "$v" in
    k1) export RESNMTF_FUSED_KIND=1;;
    k2) export RESNMTF_FUSED_KIND=2;;
    auto) ;;
    *@1) export RESNMTF_FUSED_KIND=1 RESNMTF_B200_LIB=$PWD/ab/lib_${v%@1}.so;;
    *) export RESNMTF_FUSED_KIND=2 RESNMTF_B200_LIB=$PWD/ab/lib_$v.so;;
  esac
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded --no-ksweep > "$o/ab_$v.json" 2> "$o/ab_$v.err"
  python - "$o/ab_$v.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1], "value", round(d["value"],1), "frac", round(r["frac"],4), "us/launch", round(r["us_per_launch"],2), {k:round(v,1) for k,v in r["isolated"]["us_per_launch_by_k"].items()}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
except Exception as e:
    print("no bench line:", e)
PY
done
unset RESNMTF_B200_LIB
if [ -f ab/lib_f2trace.so ]; then
  for k in 3 8; do
  RESNMTF_FUSED_KIND=2 RESNMTF_B200_LIB=$PWD/ab/lib_f2trace.so RESNMTF_FU_TIMELINE=1 timeout 120 python tools/profile_run.py --k $k --iters 30 > "$o/tl_k$k.txt" 2>&1
  done
  tail -32 "$o/tl_k8.txt"
fi
unset RESNMTF_FUSED_KIND
