bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_model_gpu.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01o.txt 2>&1
echo "layers exit $?"; head -1 gpurun_out/layers_r01o.txt; grep -E "64-> 768|256->  80|256->   2|128-> 128 k3 s1 @64x64 cs128|256-> 256 k3 s1 @32x32 cs256|512-> 512|->  27 k3 s1 @64|->  27 k3 s1 @32" gpurun_out/layers_r01o.txt | sort | uniq -c | sort -rn | head -14
