#!/bin/bash
# One GPU visit: bench line, per-layer table, ncu launch list, ncu --set full on the top kernels.
# usage: bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; cat gpurun_out/bench_$tag.json; tail -n 5 gpurun_out/bench_$tag.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
echo "ref bench exit $?"; cat gpurun_out/bench_ref_$tag.json
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_$tag.txt 2>&1
echo "layers exit $?"; head -n 3 gpurun_out/layers_$tag.txt
timeout 300 python tools/bench_configs.py > gpurun_out/configs_$tag.txt 2>&1
cat gpurun_out/configs_$tag.txt | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
   --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu launches exit $?"
for k in "decode:decode_:decode" "dcn64:dcn_ws:dcn64" "conv16:conv_rows:rows16" "s2d:conv_rows:stem_s2d" "conv256:conv_tma:tma256" "head:conv_tma:head3x3"; do
  what=${k%%:*}; rest=${k#*:}; pat=${rest%%:*}; name=${rest#*:}
  cnt=1; [ "$what" = decode ] && cnt=2     # decode = stream kernel + merge kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c $cnt \
     -o gpurun_out/prof_${name}_$tag -f python tools/run_one.py $what > gpurun_out/ncu_${name}_$tag.log 2>&1
  echo "ncu $name exit $?"
done
ls -la gpurun_out | tail -30
