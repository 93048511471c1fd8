timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x --timeout=300 -k "deconv" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s --timeout=300 -k "resnet" 2>&1 | grep -E "rel-L2|passed|failed|Error|error|assert" | head -40
