for v in t512i8 t512i4 t1024i4 t1024i8; do
  echo "== $v"; PSKB200_LIB=_variants/libpskb200_$v.so timeout 200 python tools/dense_bench.py --vec 2>&1 | tail -2
done
