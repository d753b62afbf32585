#!/bin/bash
TAG=${1:-r2aa}
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/${TAG}_tests.log
for W in 12 16 20 12 16 20; do
  BENCH_E2E_WORKERS=$W timeout 300 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_w$W.json 2> $OUT/${TAG}_w$W.err
  python - <<PY
import json
b = json.loads(open("$OUT/${TAG}_w$W.json").read().strip().splitlines()[-1])
print("workers $W value %.4g e2e %.4g (persistent %.4g) ms/step %.3f conv solo %.4f" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], b["roofline"]["stage_ms_solo_batch"]["conv"]))
PY
done
