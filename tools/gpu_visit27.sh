timeout 900 python tools/bench_configs.py 2>&1 | tail -8
