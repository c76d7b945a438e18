timeout 900 python -m pytest tests/test_abi_eval.py tests/test_golden.py tests/test_api.py -m gpu -x -q 2>&1 | tail -3
for v in "" "--threads 256" "--threads 384"; do timeout 300 python tools/phase_profile.py --no-phase $v 2>&1 | grep -E "merge_eval avg|^threads|rror"; done
timeout 300 python tools/conv_bench.py 2>&1 | grep -E "per-conv|conv_3d 512"
