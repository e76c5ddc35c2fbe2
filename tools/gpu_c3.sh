#!/usr/bin/env bash
# C3 structure on one GPU under the kernel-selection switches (one line per variant).
o=gpurun_out/c3; mkdir -p $o
python tools/small_bench.py --only c3 > $o/default.log 2>&1; cat $o/default.log
RESNMTF_NO_PDL=1 python tools/small_bench.py --only c3 > $o/nopdl.log 2>&1; cat $o/nopdl.log
RESNMTF_FUSED_KIND=2 python tools/small_bench.py --only c3 > $o/kind2.log 2>&1; cat $o/kind2.log
RESNMTF_IMPL=3 python tools/small_bench.py --only c3 > $o/twopass.log 2>&1; cat $o/twopass.log
RESNMTF_FU_TIMELINE=1 python tools/small_bench.py --only c3 > $o/timeline.log 2>&1; grep -v "^$" $o/timeline.log | head -60
