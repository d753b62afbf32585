#!/bin/bash
TAG=${1:-r2ab}
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --workload sharded --shard-as 0/8 --no-cpu-baseline --sustained-seconds 0 > $OUT/${TAG}_shard08.json 2> $OUT/${TAG}_shard08.err; echo "shard 0/8 rc=$?"; tail -2 $OUT/${TAG}_shard08.err
python - <<PY
import json
b = json.loads(open("$OUT/${TAG}_shard08.json").read().strip().splitlines()[-1])
print("shard 0/8: value %.4g e2e %.4g parity %s" % (b["value"], b["e2e"]["value"], json.dumps(b["parity"])))
PY
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
python - <<PY
import json
b = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g (persistent %.4g) ms/step %.3f parity %s" % (b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], b.get("parity")))
print("sustained %.4g" % b["sustained"]["value"], "roofline", b["roofline"]["achieved"], b["roofline"]["frac"], b["roofline"]["avg_launch_ms"])
for k, v in (b.get("other_configs") or {}).items():
    print("other", k, "value %.4g e2e %.4g ms %.2f parity %s" % (v["value"], v["e2e"]["value"], v["ms_per_step"], (v.get("parity") or {}).get("ok")))
PY
