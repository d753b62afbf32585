#!/bin/bash
# N-GPU visit: the two-engines-in-one-process test, then the bench under torchrun (default line incl. other configs).
# usage (under gpurun --gpus N): bash tools/gpu_multi2.sh <tag> <N>
TAG=${1:-r2n2}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/${TAG}_gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "two_engines or concurrent_host" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench N=$N rc=$?"; tail -3 $OUT/${TAG}_bench.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("N=%d value %.4g e2e %.4g (persistent %.4g) ms/step %.3f ranks %s parity %s" % (b["n_gpus"], b["value"], b["e2e"]["value"], b["e2e"]["persistent"]["value"], b["ms_per_step"], b["ms_per_step_ranks"], (b.get("parity") or {}).get("bases_identical")))
    print("sustained", b.get("sustained"))
    for k, v in (b.get("other_configs") or {}).items():
        print("other", k, "value %.4g e2e %.4g ms %.2f ranks %s parity %s" % (v["value"], v["e2e"]["value"], v["ms_per_step"], v.get("ms_per_step_ranks"), (v.get("parity") or {}).get("bases_identical")))
except Exception as e:
    print("no bench line", e)
PY
cat $OUT/${TAG}_bench_ref.json | cut -c1-300
