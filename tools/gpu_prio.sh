#!/bin/bash
# A/B: launch priority of the scan kernels (SCRAPPIE_B200_SCAN_PRIO)
TAG=${1:-r2ad}
OUT=gpurun_out
for i in 1 2 3; do for P in 0 1; do
  SCRAPPIE_B200_SCAN_PRIO=$P timeout 300 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_p${P}_$i.json 2> $OUT/${TAG}_p${P}_$i.err
  python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_p${P}_$i.json").read().strip().splitlines()[-1])
    c = b["roofline"]["stage_ms_per_batch_concurrent"]
    print("prio $P run $i value %.4g e2e %.4g (persistent %.4g) ms/step %.3f scan1 conc %.3f head conc %.3f parity %s" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], c["scan1"], c["head_gemm"], b["parity"]["ok"]))
except Exception as e:
    print("prio $P run $i failed", e); print(open("$OUT/${TAG}_p${P}_$i.err").read()[-600:])
PY
done; done
