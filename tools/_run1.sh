timeout 900 python -m pytest tests/test_abi_eval.py tests/test_golden.py tests/test_api.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/conv_bench.py 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1
