#!/bin/bash
# Evidence visit of round 2: parity tests, smoke, the default bench line (all configs) and the reference arm, the ncu
# launch list of the same command, ncu --set full captures of the main kernels, compute-sanitizer.
# usage (under gpurun): bash tools/gpu_round2.sh <tag> [tests|notests] [prof|noprof] [san|nosan]
TAG=${1:-r2x}
DO_TESTS=${2:-tests}
DO_PROF=${3:-prof}
DO_SAN=${4:-san}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$DO_TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
  tail -3 $OUT/${TAG}_tests.log
  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
fi
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"; cat $OUT/${TAG}_bench_ref.json | cut -c1-400
if [ "$DO_PROF" = "prof" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
  for K in ${KERNELS:-gru_scan_kernel affine_tc head_softmax decode_transducer_warp conv_act}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/${TAG}_${K} \
        python tools/prof_one.py > $OUT/${TAG}_ncu_${K}.log 2>&1; echo "ncu $K rc=$?"
  done
fi
if [ "$DO_SAN" = "san" ]; then
  bash tools/gpu_sanitize.sh $TAG
fi
python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], b.get("parity")))
    print("sustained", b.get("sustained"))
    for k, v in (b.get("other_configs") or {}).items():
        print("other", k, "value %.4g e2e %.4g ms %.2f parity %s" % (v["value"], v["e2e"]["value"], v["ms_per_step"], (v.get("parity") or {}).get("bases_identical")))
    print("cpu", b.get("cpu_baseline"))
except Exception as e:
    print("no bench line", e)
PY
ls -la $OUT | grep $TAG | head -40
