bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py
python tools/rows_bench.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01k.txt 2>&1
echo "layers exit $?"; head -3 gpurun_out/layers_r01k.txt; grep -E "^dcn" gpurun_out/layers_r01k.txt | head -16; grep -E "64-> 768|256->  80|128-> 128 k3 s1 @64x64 cs128|256-> 256 k3 s1 @32x32 cs256" gpurun_out/layers_r01k.txt | head -6
timeout 120 python tools/decode_bench.py
