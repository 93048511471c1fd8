bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_model_gpu.py
python tools/rows_bench.py
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01m.json 2> gpurun_out/bench_r01m.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r01m.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_decode']['ms'])"; tail -n 3 gpurun_out/bench_r01m.err
