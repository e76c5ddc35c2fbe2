#!/usr/bin/env bash
# Final evidence of round 2 on one box: GPU test suite, the bench line, the ncu launch list of the same command, and
# ncu --set full of the two-tile TMA kernels (k = 12).
set -u
o=gpurun_out/r02final; mkdir -p "$o"
python -m pytest tests -m gpu -x -q > "$o/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$o/pytest_gpu.log"; tail -3 "$o/pytest_gpu.log"
timeout 600 python bench.py > "$o/bench_1gpu.json" 2> "$o/bench_1gpu.err"; echo "bench rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$o/launches.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded --no-ksweep > "$o/ncu_launch.log" 2>&1; echo "ncu launches rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rn_f_step_tma -s 2 -c 1 -o "$o/f_tma_k12" -f python tools/profile_run.py --k 12 --iters 4 > "$o/ncu_f12.log" 2>&1; echo "ncu f12 rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rn_g_step_tma -s 2 -c 1 -o "$o/g_tma_k12" -f python tools/profile_run.py --k 12 --iters 4 > "$o/ncu_g12.log" 2>&1; echo "ncu g12 rc $?"
ls -la "$o"
