cd "$(dirname "$0")/../.."
for rep in 1 2; do for st in 0 4 5; do UMV_OPROJ_STAGES=$st python tools/decode_time.py 2>/dev/null | tail -1 | cut -c1-120; done; done
