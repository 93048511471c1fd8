bash tools_gpu_tests.sh tests/test_decode_gpu.py tests/test_conv_gpu.py
timeout 120 python tools/decode_bench.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01e.txt 2>&1
echo "layers exit $?"; grep -E "^B=|dcn" gpurun_out/layers_r01e.txt | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_scan -s 2 -c 1 \
   -o gpurun_out/prof_decode_r01f -f python tools/run_one.py decode > gpurun_out/ncu_decode_r01f.log 2>&1
echo "ncu decode exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_ws -s 2 -c 1 \
   -o gpurun_out/prof_dcn64_r01f -f python tools/run_one.py dcn64 > gpurun_out/ncu_dcn64_r01f.log 2>&1
echo "ncu dcn exit $?"
