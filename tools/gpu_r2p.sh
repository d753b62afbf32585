#!/bin/bash
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "caller or pool or short_read or concurrent" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
for W in 8 12 16 24; do
  BENCH_E2E_WORKERS=$W timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench_w$W.json 2> $OUT/${TAG}_bench_w$W.err; echo "bench w$W rc=$?"
done
for W in 16 32; do
  BENCH_E2E_WORKERS=$W timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --workload mixed --sets 6 --steps 12 --warmup 6 > $OUT/${TAG}_bench_mixed_w$W.json 2> $OUT/${TAG}_bench_mixed_w$W.err; echo "mixed w$W rc=$?"
done
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
tail -5 $OUT/${TAG}_bench_w8.err
