#!/bin/bash
# last visit: full GPU suite, smoke, default bench line, determinism hunts (short)
mkdir -p gpurun_out
o=gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "^$" | tail -n 40 > $o/r02s3zz.tests.txt; tail -n 8 $o/r02s3zz.tests.txt
ls $o/variant_first_failure_* 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r02s3zz_bench.json 2> $o/r02s3zz_bench.err; echo "bench exit $?"; cut -c1-260 $o/r02s3zz_bench.json; tail -n 2 $o/r02s3zz_bench.err
timeout 300 python tools/conv_race_hunt.py 800 > $o/r02s3zz_determinism.txt 2>&1
timeout 300 python tools/net_race_hunt.py 300 >> $o/r02s3zz_determinism.txt 2>&1
cat $o/r02s3zz_determinism.txt
