#!/usr/bin/env bash
# Round-2 profile pass on one box: cluster shapes, launch list of the bench command, ncu --set full of both one-pass
# kernel generations at k = 8, per-row-group trace of the second generation.
set -u
o=gpurun_out/r02prof; mkdir -p "$o"
./tools/cluster_probe > "$o/cluster_probe.txt" 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded --no-ksweep > "$o/bench_short.json" 2> "$o/bench_short.err"; echo "bench rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$o/launches.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded --no-ksweep > "$o/ncu_launch.log" 2>&1; echo "ncu launches rc $?"
timeout 120 python tools/profile_run.py --k 8 --iters 6 > "$o/run_k1.txt" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rn_fused_step -s 3 -c 1 -o "$o/fused1_k8" -f python tools/profile_run.py --k 8 --iters 6 > "$o/ncu_f1.log" 2>&1; echo "ncu f1 rc $?"
RESNMTF_FUSED_KIND=2 timeout 120 python tools/profile_run.py --k 8 --iters 6 > "$o/run_k2.txt" 2>&1
RESNMTF_FUSED_KIND=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rn_fused2_step -s 3 -c 1 -o "$o/fused2_k8" -f python tools/profile_run.py --k 8 --iters 6 > "$o/ncu_f2.log" 2>&1; echo "ncu f2 rc $?"
for k in 3 8; do
RESNMTF_FUSED_KIND=2 RESNMTF_B200_LIB=$PWD/ab/lib_f2trace.so RESNMTF_FU_TIMELINE=1 timeout 120 python tools/profile_run.py --k $k --iters 30 > "$o/tl2_k$k.txt" 2>&1
RESNMTF_FU_TIMELINE=1 timeout 120 python tools/profile_run.py --k $k --iters 30 > "$o/tl1_k$k.txt" 2>&1
done
cat "$o/run_k1.txt" "$o/run_k2.txt"; ls -la "$o"
