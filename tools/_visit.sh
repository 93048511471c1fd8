bash tools_gpu_tests.sh tests/test_conv_gpu.py tests/test_model_gpu.py
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01q.json 2> gpurun_out/bench_r01q.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r01q.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_decode']['ms'])"; tail -n 3 gpurun_out/bench_r01q.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
   --log-file gpurun_out/launches_r01q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench_r01q.log 2>&1
echo "ncu launches exit $?"
