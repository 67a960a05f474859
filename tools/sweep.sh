for c in "X=1" "UMV_SPLITS_RES=4"; do echo "== $c"; env $c timeout 200 python tools/decode_trace.py 2>&1 | grep -A12 "per kernel class" | tail -8; done
