#!/usr/bin/env python
"""Print the scan kernel's per-step hand-over timeline, every read group of CTA 0 (SCRAPPIE_B200_TRACE=1; GPU only).
usage: python tools/scan_trace.py [nreads] [model]"""
import ctypes as C
import os
import sys

import numpy as np

os.environ["SCRAPPIE_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb  # noqa: E402
from scrappie_b200.synthetic import synthetic_read  # noqa: E402

nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = sys.argv[2] if len(sys.argv) > 2 else "rgrgr_r94"
eng = sb.Engine(0)
sigs = [synthetic_read(1000 + i, 4000 if model != "rnnrf_r94" else 1000) for i in range(nreads)]
b = eng.batch(model, [len(s) for s in sigs])
b.upload(sigs)
b.forward()
b.sync()
out = np.zeros(512, dtype=np.int64)
L = sb.lib()
L.sb2_engine_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert L.sb2_engine_read_trace(eng._h, out.ctypes.data, 512) == 0
names = ["I:bar_h woke", "I:r,z issued", "I:store+fill done", "I:bar_rh woke", "I:c issued", "G:bar_x woke", "G:bar_r woke",
         "G:rh arrived", "G:bar_z woke", "G:z done", "G:bar_c woke", "G:h arrived"]
t = out.reshape(-1, 4, 16)                  # [group][step - 100][slot]
ngrp = int((t[:, 0, 0] != 0).sum())
base = t[:ngrp, 0, 0].min()
print("%d reads, %s: %d groups in CTA 0; cycles relative to the earliest group's step 100" % (nreads, model, ngrp))
for g in range(ngrp):
    print("group %d: step length %s" % (g, [int(t[g, s + 1, 0] - t[g, s, 0]) for s in range(3)]))
    ev = sorted([(int(t[g, 1, i] - base), names[i]) for i in range(12)])
    t0 = int(t[g, 1, 0] - base)
    for dt, nm in ev:
        print("   +%6d (%+5d)  %s" % (dt, dt - t0, nm))
