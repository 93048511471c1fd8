mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_ws -s 2 -c 1 \
   -o gpurun_out/prof_dcn64_r01c -f python tools/run_one.py dcn64 > gpurun_out/ncu_dcn64_r01c.log 2>&1
echo "ncu dcn64 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 2 -c 1 \
   -o gpurun_out/prof_conv64_r01c -f python tools/run_one.py conv64 > gpurun_out/ncu_conv64_r01c.log 2>&1
echo "ncu conv64 exit $?"
