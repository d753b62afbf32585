#!/bin/bash
# quick GPU visit: parity tests + bench line.  usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
else
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"
fi
tail -15 $OUT/${TAG}_tests.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    b=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g ms/step %.3f"%(b["value"],b["e2e"]["value"],b["ms_per_step"]))
    print({k:round(v,3) for k,v in (b["roofline"].get("stage_ms_solo_batch") or b["roofline"]["stage_ms_per_batch"]).items()})
except Exception as e:
    print("no bench line", e)
PY
