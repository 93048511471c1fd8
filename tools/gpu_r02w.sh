#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 8 --error-exitcode 0 python -m pytest "tests/test_conv_gpu.py" -k "dcnv2_matches_torchvision and 2-64-64-24-32" -q -m gpu -x --tb=line > gpurun_out/r02w.race_fp.txt 2>&1
grep -c "Race reported" gpurun_out/r02w.race_fp.txt
grep -v "^=========     at\|^=========     by\|^=========         in\|^$" gpurun_out/r02w.race_fp.txt | head -60
