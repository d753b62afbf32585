#!/bin/bash
TAG=${1:-r2o}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --workload mixed --sets 6 --steps 12 --warmup 6 > $OUT/${TAG}_bench_mixed.json 2> $OUT/${TAG}_bench_mixed.err; echo "mixed rc=$?"
BENCH_E2E_WORKERS=32 timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --workload mixed --sets 6 --steps 12 --warmup 6 > $OUT/${TAG}_bench_mixed_w32.json 2> $OUT/${TAG}_bench_mixed_w32.err; echo "mixed w32 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s allocs %s" % (
            b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], (b.get("parity") or {}).get("bases_identical"), b["e2e"].get("workspace_allocations_in_timed_region")))
    except Exception as e:
        print(f, "no bench line", e)
PY
