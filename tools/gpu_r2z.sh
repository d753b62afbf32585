#!/bin/bash
TAG=${1:-r2z}
OUT=gpurun_out
for W in 48 64; do
  BENCH_MIXED_WORKERS=$W timeout 800 python bench.py --no-cpu-baseline --sustained-seconds 0 > $OUT/${TAG}_bench_w$W.json 2> $OUT/${TAG}_bench_w$W.err; echo "rc=$?"; tail -3 $OUT/${TAG}_bench_w$W.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_w*.json")):
    b = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"]))
    for k, v in (b.get("other_configs") or {}).items():
        print("other", k, "value %.4g e2e %.4g (persistent %.4g) ms %.2f allocs %s parity %s %s" % (v["value"], v["e2e"]["value"], v["e2e"]["persistent"], v["ms_per_step"], v["e2e"]["workspace_allocations_in_timed_region"], (v.get("parity") or {}).get("bases_identical"), v["e2e"]["api"][-50:]))
PY
