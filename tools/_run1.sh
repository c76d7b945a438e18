timeout 900 python -m pytest tests/test_abi_eval.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/phase_profile.py --no-phase 2>&1 | grep -E "merge_eval avg|rror"
timeout 300 python tools/phase_profile.py 2>&1 | tail -9
