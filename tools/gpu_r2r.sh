#!/bin/bash
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3; do
  timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_s1_$i.json 2> $OUT/${TAG}_s1_$i.err; echo "steps1 run $i rc=$?"; tail -2 $OUT/${TAG}_s1_$i.err
done
MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_mc.json 2> $OUT/${TAG}_mc.err; echo "malloc_check rc=$?"; tail -3 $OUT/${TAG}_mc.err
MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_mc4.json 2> $OUT/${TAG}_mc4.err; echo "malloc_check steps4 rc=$?"; tail -3 $OUT/${TAG}_mc4.err
for W in 16 24; do
  BENCH_E2E_WORKERS=$W timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench_w$W.json 2> $OUT/${TAG}_bench_w$W.err; echo "bench w$W rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s allocs %s" % (
            b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], (b.get("parity") or {}).get("bases_identical"), b["e2e"].get("workspace_allocations_in_timed_region")))
    except Exception as e:
        print(f, "no bench line", e)
PY
