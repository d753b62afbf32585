#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "layerwise or ragged or batch_size or full_size" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_tests.log | tail -5
for W in 16 8 24; do
  BENCH_E2E_WORKERS=$W timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench_w$W.json 2> $OUT/${TAG}_bench_w$W.err; echo "bench w$W rc=$?"
done
for W in 16 8; do
  BENCH_E2E_WORKERS=$W timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --model rnnrf_r94 --steps 8 --warmup 4 > $OUT/${TAG}_bench_rnnrf_w$W.json 2>> $OUT/${TAG}_bench_w$W.err; echo "rnnrf w$W rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s" % (
            b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], (b.get("parity") or {}).get("bases_identical")))
        r = b["roofline"]
        print("   solo", {k: round(v, 3) for k, v in r["stage_ms_solo_batch"].items()})
        print("   conc", {k: round(v, 3) for k, v in r["stage_ms_per_batch_concurrent"].items()})
    except Exception as e:
        print(f, "no bench line", e)
PY
