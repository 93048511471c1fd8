bash tools_gpu_tests.sh tests/test_conv_gpu.py
python tools/rows_bench.py
