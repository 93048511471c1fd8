echo "=== fp32 blend"; timeout 600 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q -x -s --timeout=300 -k "oracle or dcn" 2>&1 | grep -E "rel-L2|passed|failed|Error" | head -30
echo "=== bf16 blend"; CNB_DCN_BLEND=bf16 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q -x -s --timeout=300 -k "oracle or dcn" 2>&1 | grep -E "rel-L2|passed|failed|Error" | head -30
CNB_DCN_BLEND=bf16 timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01l.txt 2>&1
head -1 gpurun_out/layers_r01l.txt; grep -E "^dcn" gpurun_out/layers_r01l.txt | head -16
