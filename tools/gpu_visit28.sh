timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 2 -c 1 \
   -o gpurun_out/prof_head1x1_r01n -f python tools/run_one.py head1x1 > gpurun_out/ncu_head1x1.log 2>&1
echo "ncu head1x1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 2 -c 1 \
   -o gpurun_out/prof_conv128_r01n -f python tools/run_one.py conv128 > gpurun_out/ncu_conv128.log 2>&1
echo "ncu conv128 exit $?"
