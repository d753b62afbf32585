#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list, ncu full capture of the scan kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests] [ncu-kernel-regex]
TAG=${1:-rX}
DO_TESTS=${2:-tests}
KREGEX=${3:-gru_scan}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$DO_TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?" 
  tail -3 $OUT/${TAG}_tests.log
  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 6 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -12
