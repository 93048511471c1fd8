#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r02r.txt; : > $o
echo "== warm" >> $o
timeout 300 python tools/dcn_bench.py >> $o 2>&1
echo "== cold" >> $o
DCN_BENCH_COLD=1 timeout 300 python tools/dcn_bench.py >> $o 2>&1
echo "== cold REACH=1" >> $o
DCN_BENCH_COLD=1 CNB_DCN_REACH=1 timeout 300 python tools/dcn_bench.py d64 d128 d256 >> $o 2>&1
cat $o
