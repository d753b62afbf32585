#!/bin/bash
# compute-sanitizer over tools/sanitize_smoke.py: memcheck (out-of-bounds / misaligned accesses), synccheck (barrier
# misuse), initcheck (reads of uninitialised global memory).  initcheck does not see writes made by the TMA engine
# (cp.async.bulk shared -> global): every buffer the tcgen05 scan stores its results with looks "uninitialised" to it,
# so by default it runs with SCRAPPIE_B200_SCAN=ffma (plain stores everywhere: must be clean); `initcheck` as shipped is
# opt-in (tens of thousands of reports on reads of scan outputs: > 15 min).
# usage (under gpurun): bash tools/gpu_sanitize.sh <tag> [memcheck synccheck initcheck_ffma initcheck]
TAG=${1:-r2}
shift
TOOLS=${@:-memcheck synccheck initcheck_ffma}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/sanitize_smoke.py > $OUT/${TAG}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -1 $OUT/${TAG}_sanitize_plain.log
for T in $TOOLS; do
  TOOL=${T%%_*}
  if [ "$T" = "initcheck_ffma" ]; then export SCRAPPIE_B200_SCAN=ffma; else unset SCRAPPIE_B200_SCAN; fi
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 40 --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_sanitize_$T.log 2>&1
  echo "$T rc=$?"; grep -E "ERROR SUMMARY|sanitize smoke ok" $OUT/${TAG}_sanitize_$T.log | head -3
  grep -E "Device Frame" $OUT/${TAG}_sanitize_$T.log | sed 's/+0x.*//' | sort | uniq -c | head -6
done
