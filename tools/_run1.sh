timeout 900 python -m pytest tests/test_abi_eval.py tests/test_golden.py tests/test_api.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/conv_bench.py 2>&1 | grep -E "per-conv|conv_3d 512"
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_line.json; cat gpurun_out/bench_line.json | cut -c1-150
timeout 300 python tools/c1_latency.py 2>&1 | tail -3
