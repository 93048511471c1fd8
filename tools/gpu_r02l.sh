#!/bin/bash
# ncu full capture of the footprint DCN kernel (64->64 @128x128, B=32)
mkdir -p gpurun_out
o=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_fp -s 2 -c 1 -o $o/prof_dcnfp64_r02l -f python tools/run_one.py dcn64 > $o/ncu_dcnfp64_r02l.log 2>&1
echo "ncu exit $?"; tail -n 3 $o/ncu_dcnfp64_r02l.log
