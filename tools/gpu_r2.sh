#!/bin/bash
# Round-2 GPU visit: parity tests (no -x: every failure is wanted), smoke, A/B of the scan generations.
# usage (under gpurun): bash tools/gpu_r2.sh <tag> [pytest -k expr]
TAG=${1:-r2}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -k "$KEXPR" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
else
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
fi
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_tests.log | tail -20
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
for GEN in 5 4; do
  SCRAPPIE_B200_SCAN_GEN=$GEN timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_gen$GEN.json 2> $OUT/${TAG}_bench_gen$GEN.err; echo "bench gen$GEN rc=$?"
  SCRAPPIE_B200_SCAN_GEN=$GEN timeout 600 python bench.py --no-cpu-baseline --model rnnrf_r94 --steps 8 --warmup 4 > $OUT/${TAG}_bench_rnnrf_gen$GEN.json 2>> $OUT/${TAG}_bench_gen$GEN.err; echo "rnnrf gen$GEN rc=$?"
done
python - <<PY
import json
for name in ("bench_gen5", "bench_gen4", "bench_rnnrf_gen5", "bench_rnnrf_gen4"):
    try:
        b = json.loads(open("$OUT/${TAG}_%s.json" % name).read().strip().splitlines()[-1])
        print(name, "value %.4g e2e %.4g ms/step %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"]))
        print("   solo", {k: round(v, 3) for k, v in b["roofline"]["stage_ms_solo_batch"].items()})
        print("   conc", {k: round(v, 3) for k, v in b["roofline"]["stage_ms_per_batch_concurrent"].items()})
    except Exception as e:
        print(name, "no bench line", e)
PY
