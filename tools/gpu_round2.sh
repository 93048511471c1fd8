#!/bin/bash
# One GPU visit of round 2: bench lines (default, train, other configs), per-layer table, training kernel table,
# ncu launch list, ncu --set full on the top / new kernels.   usage: bash tools/gpu_round2.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
o=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > $o/bench_$tag.json 2> $o/bench_$tag.err
echo "bench exit $?"; cut -c1-400 $o/bench_$tag.json; tail -n 3 $o/bench_$tag.err
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > $o/bench_train_$tag.json 2>> $o/bench_$tag.err; echo "train exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_ref_$tag.json 2>> $o/bench_$tag.err
for c in 1 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $o/bench_config${c}_$tag.json 2>> $o/bench_$tag.err; echo "config $c exit $?"
done
timeout 300 python tools/profile_layers.py 32 > $o/layers_$tag.txt 2>&1; head -n 2 $o/layers_$tag.txt
timeout 600 python tools/profile_train.py 2 > $o/train_kernels_$tag.txt 2>&1; head -n 4 $o/train_kernels_$tag.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
   --log-file $o/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train-leg > $o/ncu_bench_$tag.log 2>&1
echo "ncu launches exit $?"
# DRAM traffic of the tensor-core family over one eager step (two cheap metrics, one pass): -> profiles/traffic.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_|dcn_|head_fused" -c 400 --csv \
   --log-file $o/traffic_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-train-leg > $o/ncu_traffic_$tag.log 2>&1
echo "ncu traffic exit $?"
for k in "headfused:head_fused:headfused" "wgrad64:conv_wgrad:wgrad64" "wgrad256:conv_wgrad:wgrad256" "col2im64:dcn_col2im:col2im64" "im2col64:dcn_im2col:im2col64" "dcn64:dcn_fp:dcn64" "decode:decode_:decode" "conv256:conv_tma:tma256"; do
  what=${k%%:*}; rest=${k#*:}; pat=${rest%%:*}; name=${rest#*:}
  cnt=1; [ "$what" = decode ] && cnt=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c $cnt \
     -o $o/prof_${name}_$tag -f python tools/run_one.py $what > $o/ncu_${name}_$tag.log 2>&1
  echo "ncu $name exit $?"
done
ls -la $o | grep $tag | tail -n 40
