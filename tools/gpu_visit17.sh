bash tools_gpu_tests.sh tests/test_conv_gpu.py
python tools/rows_bench.py
CNB_ROWS_R=2 python tools/rows_bench.py c16 stem
