bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_decode_gpu.py tests/test_losses_gpu.py tests/test_model_gpu.py
bash tools/gpu_round.sh r01b
