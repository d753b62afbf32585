#!/bin/bash
TAG=${1:-r2w}
OUT=gpurun_out
mkdir -p $OUT
for W in 48 64; do
BENCH_E2E_WORKERS=$W timeout 600 python bench.py --workload mixed --steps 12 --warmup 6 --sets 6 --no-cpu-baseline --sustained-seconds 0 > $OUT/${TAG}_mixed_w$W.json 2> $OUT/${TAG}_mixed_w$W.err; echo "mixed w$W rc=$?"; tail -2 $OUT/${TAG}_mixed_w$W.err | cut -c1-300
done
for S in 4 6 8; do
  timeout 300 python bench.py --sets $S --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_fixed_s$S.json 2> $OUT/${TAG}_fixed_s$S.err; echo "fixed sets $S rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f e2e-ms %.2f allocs %s" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["e2e"].get("workspace_allocations_in_timed_region")))
    except Exception as e:
        print(f, "no bench line", e)
PY
