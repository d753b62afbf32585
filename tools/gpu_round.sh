#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list, ncu full captures of the main kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-rX}
DO_TESTS=${2:-tests}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$DO_TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
  tail -3 $OUT/${TAG}_tests.log
  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"; cat $OUT/${TAG}_bench_ref.json
timeout 600 python bench.py --model rnnrf_r94 --no-cpu-baseline > $OUT/${TAG}_bench_rnnrf.json 2>> $OUT/${TAG}_bench.err; echo "rnnrf rc=$?"; cat $OUT/${TAG}_bench_rnnrf.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
for K in ${KERNELS:-gru_scan_v4 decode_transducer_v2 head_softmax affine_tc conv_act}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/${TAG}_${K} \
      python tools/prof_one.py > $OUT/${TAG}_ncu_${K}.log 2>&1; echo "ncu $K rc=$?"
done
ls -la $OUT | tail -14
