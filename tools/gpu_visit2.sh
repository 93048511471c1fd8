bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_model_gpu.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01c.txt 2>&1
echo "layers exit $?"; grep -E "^B=|dcn" gpurun_out/layers_r01c.txt | head -40
