#!/usr/bin/env python
"""Small driver for ncu captures: one batch (default 256 x 4000-sample reads), forward + decode, `reps` times.
    ncu --set full -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/x python tools/prof_one.py [model] [nread] [nsample] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb
from scrappie_b200.synthetic import synthetic_read

model = sys.argv[1] if len(sys.argv) > 1 else "rgrgr_r94"
nread = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nsample = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
eng = sb.Engine(0)
sigs = [synthetic_read(1000 + i, nsample) for i in range(nread)]
b = eng.batch(model, [len(s) for s in sigs])
b.upload(sigs)
for _ in range(reps):
    b.forward()
    b.decode()
b.sync()
tot, fwd, dec = b.time(nrep=3)
print("ms total %s forward %s decode %s" % (tot, fwd, dec))
print({k: round(v, 4) for k, v in b.stage_ms().items()})
