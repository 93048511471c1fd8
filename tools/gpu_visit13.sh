bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_decode_gpu.py
for k in uniform; do timeout 120 python tools/decode_timeline.py $k; done
timeout 120 python tools/decode_bench.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01j.txt 2>&1
echo "layers exit $?"; head -12 gpurun_out/layers_r01j.txt; grep -E "64->  27 k3 s1 @128|64-> 768" gpurun_out/layers_r01j.txt | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 2 -c 1 \
   -o gpurun_out/prof_rows16_r01j -f python tools/run_one.py conv16 > gpurun_out/ncu_rows16_r01j.log 2>&1
echo "ncu rows16 exit $?"
