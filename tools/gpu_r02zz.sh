#!/bin/bash
# comprehensive round-2 visit: full GPU suite, bench lines, profiles, sanitizers (ordered by priority)
mkdir -p gpurun_out
o=gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "^$" | tail -n 60 > $o/r02zz.tests.txt; tail -n 25 $o/r02zz.tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -n 2
bash tools/gpu_round2.sh r02zz 2>&1 | tail -n 60
SAN="compute-sanitizer --print-limit 5 --error-exitcode 0"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool" > $o/r02zz.san_$tool.txt
  for sel in "tests/test_conv_gpu.py -k fused_center_head" "tests/test_train_gpu.py -k test_conv_forward_backward" "tests/test_train_gpu.py -k dcn_forward" "tests/test_conv_gpu.py -k dcn" "tests/test_decode_gpu.py -k golden" "tests/test_conv_gpu.py -k rows"; do
    echo "--- pytest $sel" >> $o/r02zz.san_$tool.txt
    timeout 900 $SAN --tool $tool python -m pytest $sel -q -m gpu -x --tb=line 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error:" | head -n 12 >> $o/r02zz.san_$tool.txt
  done
done
cat $o/r02zz.san_memcheck.txt $o/r02zz.san_racecheck.txt
