timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 2 -c 1 \
   -o gpurun_out/prof_rows16_r01k -f python tools/rows_bench.py c16 > gpurun_out/ncu_rows16_r01k.log 2>&1
echo "ncu rows16 exit $?"
