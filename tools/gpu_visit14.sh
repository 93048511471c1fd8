bash tools_gpu_tests.sh tests/test_conv_gpu.py
python tools/rows_bench.py
CNB_ROWS_ACC=2 python tools/rows_bench.py c16 c64 stem
CNB_ROWS_ACC=4 python tools/rows_bench.py c16 c64 stem
CNB_ROWS_DEPTH=2 python tools/rows_bench.py c16 c64 stem
CNB_ROWS_DEPTH=4 python tools/rows_bench.py c16 c64 stem
