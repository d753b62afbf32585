#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "layerwise or ragged or batch_size or full_size or outliers" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${TAG}_tests.log
for A in 148 96 74; do
  SCRAPPIE_B200_AFFINE_CTAS=$A timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench_a$A.json 2> $OUT/${TAG}_bench_a$A.err; echo "bench a$A rc=$?"
done
SCRAPPIE_B200_TIMING=1 timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --workload mixed --sets 6 --steps 12 --warmup 6 > $OUT/${TAG}_bench_mixed.json 2> $OUT/${TAG}_bench_mixed.err; echo "mixed rc=$?"
tail -30 $OUT/${TAG}_bench_mixed.err | cut -c1-200
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s allocs %s" % (
            b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], (b.get("parity") or {}).get("bases_identical"), b["e2e"].get("workspace_allocations_in_timed_region")))
        r = b["roofline"]
        print("   solo", {k: round(v, 3) for k, v in r["stage_ms_solo_batch"].items()})
        print("   conc", {k: round(v, 3) for k, v in r["stage_ms_per_batch_concurrent"].items()})
    except Exception as e:
        print(f, "no bench line", e)
PY
