#!/bin/bash
# A/B of two builds of the library on the fixed and the mixed workload.  usage: bash tools/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
for V in main alt; do
  if [ $V = alt ]; then export SCRAPPIE_B200_LIB=$PWD/scrappie_b200/libscrappie_b200_alt.so; else unset SCRAPPIE_B200_LIB; fi
  timeout 600 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_bench_$V.json 2> $OUT/${TAG}_bench_$V.err; echo "bench $V rc=$?"
  timeout 600 python tools/mixed_probe2.py > $OUT/${TAG}_probe2_$V.log 2>&1; echo "probe $V rc=$?"; tail -4 $OUT/${TAG}_probe2_$V.log | cut -c1-420
done
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
