PSKB200_PRINT_CTRL3=1 PSKB200_JIT_DEFS="-DPSK_COUNT_DIRECT" timeout 200 python tools/sweep_geom.py --reps 3 --geoms 256x4096x2,256x4096x1 2>&1 | tail -12 | cut -c1-100
