bash tools_gpu_tests.sh tests/test_decode_gpu.py tests/test_conv_gpu.py
for k in uniform bumps; do timeout 120 python tools/decode_timeline.py $k; done
timeout 120 python tools/decode_bench.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01i.txt 2>&1
echo "layers exit $?"; head -12 gpurun_out/layers_r01i.txt; grep -E "64->  27 k3 s1 @128|64-> 768" gpurun_out/layers_r01i.txt | head -3
