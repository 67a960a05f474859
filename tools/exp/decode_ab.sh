#!/bin/bash
# A/B of decode-step switches, each configuration in its own process (tools/decode_time.py), alternating to cancel box drift.
cd "$(dirname "$0")/../.."
for rep in 1 2; do
  for cfg in "$@"; do
    env $cfg python tools/decode_time.py 2>/dev/null | tail -1
  done
done
