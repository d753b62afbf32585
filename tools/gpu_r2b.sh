#!/bin/bash
# Round-2 visit B: the new bench line (all configs), then CTA-cap experiments on the HBM-bound tcgen05 kernels.
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -5 $OUT/${TAG}_bench.err
for CFG in "96 148" "64 148" "148 96" "96 96" "64 64"; do
  set -- $CFG
  SCRAPPIE_B200_AFFINE_CTAS=$1 SCRAPPIE_B200_HEAD_CTAS=$2 timeout 300 python bench.py --no-other-configs --no-cpu-baseline --sustained-seconds 0 \
      > $OUT/${TAG}_bench_a$1_h$2.json 2>> $OUT/${TAG}_bench.err; echo "a$1 h$2 rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench*.json")):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s" % (
            b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], (b.get("parity") or {}).get("bases_identical")))
        r = b["roofline"]
        print("   solo", {k: round(v, 3) for k, v in r["stage_ms_solo_batch"].items() if k in ("affine2", "scan2", "head_gemm", "decode")},
              "conc", {k: round(v, 3) for k, v in r["stage_ms_per_batch_concurrent"].items() if k in ("affine2", "scan2", "head_gemm", "decode")})
        if b.get("sustained"): print("   sustained %.4g over %.1f s, clocks %s" % (b["sustained"]["value"], b["sustained"]["seconds"], b["sustained"]["clocks"]))
        for k, v in (b.get("other_configs") or {}).items():
            print("   other", k, "value %.4g e2e %.4g ms %.2f parity %s" % (v["value"], v["e2e"]["value"], v["ms_per_step"], (v.get("parity") or {}).get("bases_identical")))
        if b.get("cpu_baseline"): print("   cpu", b["cpu_baseline"])
    except Exception as e:
        print(f, "no bench line", e)
PY
