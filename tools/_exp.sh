for d in "-DPSK_COARSE_GROUP=8" "-DPSK_COARSE_GROUP=4" "-DPSK_COARSE_GROUP=2"; do
  echo "== $d"; PSKB200_JIT_DEFS=$d timeout 200 python tools/sweep_geom.py --geoms 256x4096x2 2>&1 | tail -1 | cut -c1-100
done
