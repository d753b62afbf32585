#!/bin/bash
# A/B of CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues that the streams of one context are mapped onto; default 8)
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
for C in ${CONNS:-8 32}; do
  CUDA_DEVICE_MAX_CONNECTIONS=$C timeout 600 python bench.py --no-cpu-baseline --sustained-seconds 0 > $OUT/${TAG}_bench_c$C.json 2> $OUT/${TAG}_bench_c$C.err; echo "conn $C rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_c*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"]))
        for k, v in (b.get("other_configs") or {}).items():
            print("   ", k, "value %.4g e2e %.4g (persistent %.4g) ms %.2f parity %s" % (v["value"], v["e2e"]["value"], v["e2e"]["persistent"], v["ms_per_step"], (v.get("parity") or {}).get("bases_identical")))
    except Exception as e:
        print(f, "no bench line", e)
PY
