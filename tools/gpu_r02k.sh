#!/bin/bash
# DCN footprint kernel: parity (fresh interpreters per variant), then single-layer timings
mkdir -p gpurun_out
o=gpurun_out/r02k.dcn.txt
: > $o
run() { echo "== $*" >> $o; env "$@" timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k dcn 2>&1 | tail -n 4 >> $o; }
run CNB_X=0
run CNB_DCN_REACH=1
run CNB_DCN_NB=3 CNB_DCN_STAGES=4
if grep -q "failed\|rror" $o; then cat $o; exit 1; fi
bench() { echo "== bench $*" >> $o; env "$@" timeout 300 python tools/dcn_bench.py >> $o 2>&1; }
bench CNB_DCN_IMPL=ws
bench CNB_X=0
bench CNB_DCN_REACH=1
bench CNB_DCN_REACH=3
bench CNB_DCN_BLEND=bf16
cat $o
