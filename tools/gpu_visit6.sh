bash tools_gpu_tests.sh tests/test_decode_gpu.py
timeout 120 python tools/decode_bench.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_scan -s 2 -c 1 \
   -o gpurun_out/prof_decode_r01e -f python tools/run_one.py decode > gpurun_out/ncu_decode_r01e.log 2>&1
echo "ncu decode exit $?"
