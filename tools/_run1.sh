BENCH_E2E_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 2> gpurun_out/e2e_verbose.log | tail -1 > gpurun_out/bench_line.json
timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_line2.json
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
