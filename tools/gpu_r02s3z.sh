#!/bin/bash
# final visit of the round: full GPU suite, smoke, bench lines, layer table, launch list, sanitizers on the plain 3x3 mode
mkdir -p gpurun_out
o=gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "^$" | tail -n 40 > $o/r02s3z.tests.txt; tail -n 12 $o/r02s3z.tests.txt
ls $o/variant_first_failure_* 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r02s3z_bench.json 2> $o/r02s3z_bench.err; echo "bench exit $?"; cut -c1-300 $o/r02s3z_bench.json; tail -n 2 $o/r02s3z_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/r02s3z_bench_ref.json 2>> $o/r02s3z_bench.err; echo "ref exit $?"
for c in 1 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $o/r02s3z_bench_config$c.json 2>> $o/r02s3z_bench.err; echo "config $c exit $?"
done
timeout 300 python tools/profile_layers.py 32 > $o/r02s3z_layers.txt 2>&1; head -n 2 $o/r02s3z_layers.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
   --log-file $o/r02s3z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train-leg > $o/r02s3z_ncu_bench.log 2>&1
echo "ncu launches exit $?"
SAN="compute-sanitizer --print-limit 5 --error-exitcode 0"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool" > $o/r02s3z.san_$tool.txt
  for sel in "tests/test_conv_gpu.py -k 20-128-2" "tests/test_conv_gpu.py -k dcnv2_matches"; do
    echo "--- pytest $sel" >> $o/r02s3z.san_$tool.txt
    timeout 600 $SAN --tool $tool python -m pytest $sel -q -m gpu -x --tb=line 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Race reported|Program hit|Invalid|error:" | head -n 12 | cut -c1-240 >> $o/r02s3z.san_$tool.txt
  done
done
cat $o/r02s3z.san_memcheck.txt $o/r02s3z.san_racecheck.txt
