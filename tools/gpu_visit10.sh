bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_model_gpu.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01h.txt 2>&1
echo "layers exit $?"; head -12 gpurun_out/layers_r01h.txt; grep -E "64->  27 k3 s1 @128|64-> 768" gpurun_out/layers_r01h.txt | head -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err
echo "bench exit $?"; cat gpurun_out/bench_r01h.json; tail -n 5 gpurun_out/bench_r01h.err
