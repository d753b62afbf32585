#!/bin/bash
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2; do for C in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$C timeout 300 python bench.py --no-cpu-baseline --no-other-configs --sustained-seconds 0 > $OUT/${TAG}_fixed_c${C}_$i.json 2> $OUT/${TAG}_fixed_c${C}_$i.err; echo "fixed conn $C run $i rc=$?"
done; done
timeout 900 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f sustained %s" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], (b.get("sustained") or {}).get("value")))
        for k, v in (b.get("other_configs") or {}).items():
            print("   ", k, "value %.4g e2e %.4g (persistent %.4g) ms %.2f allocs %s parity %s" % (v["value"], v["e2e"]["value"], v["e2e"]["persistent"], v["ms_per_step"], v["e2e"].get("workspace_allocations_in_timed_region"), (v.get("parity") or {}).get("bases_identical")))
    except Exception as e:
        print(f, "no bench line", e)
PY
