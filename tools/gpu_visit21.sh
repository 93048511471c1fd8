timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x --timeout=300 --durations=8 2>&1 | tail -20
