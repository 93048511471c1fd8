bash tools_gpu_tests.sh tests/test_model_gpu.py
timeout 300 python tools/profile_layers.py 32 > gpurun_out/layers_r01g.txt 2>&1
echo "layers exit $?"; cat gpurun_out/layers_r01g.txt | head -90
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err
echo "bench exit $?"; cat gpurun_out/bench_r01g.json; tail -n 5 gpurun_out/bench_r01g.err
