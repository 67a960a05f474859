#!/bin/bash
# compute-sanitizer over the tiny-dims hot path: memcheck, racecheck, synccheck, initcheck.
# Run on a GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Summaries -> gpurun_out/sanitizer_<tool>.txt (copy the tails into profiles/).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  start=$(date +%s)
  timeout ${SAN_TIMEOUT:-420} $CS --tool $tool $extra --print-limit 30 --error-exitcode 9 \
      python tools/sanitize_target.py ${SAN_WHICH:-vqa flow vae} > gpurun_out/sanitizer_$tool.txt 2>&1
  rc=$?
  echo "== $tool rc=$rc $(( $(date +%s) - start ))s" | tee -a gpurun_out/sanitizer_$tool.txt
  tail -n 6 gpurun_out/sanitizer_$tool.txt
done
