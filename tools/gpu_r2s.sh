#!/bin/bash
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-other-configs > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
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
        print("   sustained", b.get("sustained"))
    except Exception as e:
        print(f, "no bench line", e)
PY
