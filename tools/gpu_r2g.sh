#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_tests.log | tail -5
for W in 8 6 12; do
  BENCH_E2E_WORKERS=$W timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench_w$W.json 2> $OUT/${TAG}_bench_w$W.err; echo "bench w$W rc=$?"
done
for S in 6 8; do
  timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --sets $S --steps 24 > $OUT/${TAG}_bench_sets$S.json 2> $OUT/${TAG}_bench_sets$S.err; echo "bench sets$S rc=$?"
done
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --model rnnrf_r94 --steps 8 --warmup 4 > $OUT/${TAG}_bench_rnnrf.json 2>> $OUT/${TAG}_bench_w8.err; echo "rnnrf rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --workload mixed --sets 6 --steps 12 --warmup 6 > $OUT/${TAG}_bench_mixed.json 2>> $OUT/${TAG}_bench_w8.err; echo "mixed rc=$?"
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
