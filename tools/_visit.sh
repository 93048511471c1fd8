bash tools_gpu_tests.sh tests/test_decode_gpu.py
timeout 120 python tools/decode_bench.py
