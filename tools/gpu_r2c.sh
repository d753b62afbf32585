#!/bin/bash
# Round-2 GPU visit: parity tests (every failure is wanted), smoke, headline + rnnrf bench, scan timeline, mixed probe.
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_tests.log | tail -20
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 --model rnnrf_r94 --steps 8 --warmup 4 > $OUT/${TAG}_bench_rnnrf.json 2>> $OUT/${TAG}_bench.err; echo "rnnrf rc=$?"
timeout 300 python tools/scan_trace.py 256 > $OUT/${TAG}_scan_trace.log 2>&1; head -16 $OUT/${TAG}_scan_trace.log
timeout 300 python tools/debug_rnnrf.py > $OUT/${TAG}_debug_rnnrf.log 2>&1; tail -7 $OUT/${TAG}_debug_rnnrf.log | cut -c1-400
timeout 600 python tools/mixed_probe.py > $OUT/${TAG}_mixed_probe.log 2>&1; echo "probe rc=$?"; tail -9 $OUT/${TAG}_mixed_probe.log | cut -c1-600
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
