mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 2 -c 1 \
   -o gpurun_out/prof_rows16_r01h -f python tools/run_one.py conv16 > gpurun_out/ncu_rows16_r01h.log 2>&1
echo "ncu rows16 exit $?"; tail -3 gpurun_out/ncu_rows16_r01h.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 2 -c 1 \
   -o gpurun_out/prof_rows64_r01h -f python tools/run_one.py conv64 > gpurun_out/ncu_rows64_r01h.log 2>&1
echo "ncu rows64 exit $?"; tail -3 gpurun_out/ncu_rows64_r01h.log
for k in uniform bumps; do timeout 120 python tools/decode_timeline.py $k; done
