for div in 32 16; do
  echo "== div $div partition"; PSKB200_SAMPLE_DIV=$div PSKB200_PROFILE_PARTITION=1 timeout 200 python tools/sweep_geom.py --geoms 256x4096x2 2>&1 | tail -1
  echo "== div $div merge"; PSKB200_SAMPLE_DIV=$div timeout 200 python tools/sweep_geom.py --geoms 256x4096x2 2>&1 | tail -1
done
for v in mb6 mb8; do
  echo "== $v div 32 partition"; PSKB200_PROFILE_PARTITION=1 timeout 200 python tools/sweep_geom.py --lib _variants/libpskb200_$v.so --geoms 256x4096x2 2>&1 | tail -1
  echo "== $v div 16 partition"; PSKB200_SAMPLE_DIV=16 PSKB200_PROFILE_PARTITION=1 timeout 200 python tools/sweep_geom.py --lib _variants/libpskb200_$v.so --geoms 256x4096x2 2>&1 | tail -1
done
