#!/bin/bash
# One GPU visit: bench line, per-layer table, ncu launch list, ncu --set full on the top kernels.
# usage: bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; cat gpurun_out/bench_$tag.json; tail -n 5 gpurun_out/bench_$tag.err
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_$tag.txt 2>&1
echo "layers exit $?"; head -n 100 gpurun_out/layers_$tag.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
   --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_scan -s 2 -c 1 \
   -o gpurun_out/prof_decode_$tag -f python tools/run_one.py decode > gpurun_out/ncu_decode_$tag.log 2>&1
echo "ncu decode exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 \
   -o gpurun_out/prof_conv64_$tag -f python tools/run_one.py conv64 > gpurun_out/ncu_conv64_$tag.log 2>&1
echo "ncu conv64 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 \
   -o gpurun_out/prof_conv256_$tag -f python tools/run_one.py conv256 > gpurun_out/ncu_conv256_$tag.log 2>&1
echo "ncu conv256 exit $?"
ls -la gpurun_out
