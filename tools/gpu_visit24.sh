timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x --timeout=300 --durations=6 2>&1 | tail -12
python tools/rows_bench.py c16 stem c64
