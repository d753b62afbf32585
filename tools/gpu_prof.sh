#!/bin/bash
# ncu --set full capture of one kernel.  usage: bash tools/gpu_prof.sh <tag> <kernel-regex> [skip] [driver args...]
TAG=$1; KREGEX=$2; SKIP=${3:-1}; shift 3
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 1 -f -o $OUT/${TAG}_prof \
    python tools/prof_one.py "$@" > $OUT/${TAG}_prof.log 2>&1; echo "ncu rc=$?"; tail -4 $OUT/${TAG}_prof.log
