for w in 4 6 8 10 12 16 30; do echo "width $w"; PSK_CONV_WIDTH=$w timeout 300 python tools/conv_bench.py 2>&1 | grep -E "per-conv|conv_3d 512"; done
