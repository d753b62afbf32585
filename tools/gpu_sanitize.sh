#!/bin/bash
# compute-sanitizer over tools/sanitize_smoke.py: memcheck (out-of-bounds / misaligned accesses), initcheck (reads of
# uninitialised global memory), synccheck (barrier misuse).  usage (under gpurun): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/sanitize_smoke.py > $OUT/${TAG}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -1 $OUT/${TAG}_sanitize_plain.log
for TOOL in memcheck initcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|sanitize smoke ok|Error|Invalid|Uninitialized" $OUT/${TAG}_sanitize_$TOOL.log | head -12
done
