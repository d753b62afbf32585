#!/usr/bin/env python
"""Print the scan kernel's per-step hand-over timeline (needs SCRAPPIE_B200_TRACE=1; GPU only)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["SCRAPPIE_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb
from scrappie_b200.synthetic import synthetic_read

eng = sb.Engine(0)
sigs = [synthetic_read(1000 + i, 4000) for i in range(64)]
b = eng.batch("rgrgr_r94", [len(s) for s in sigs])
b.upload(sigs)
b.forward()
b.sync()
out = np.zeros(64, dtype=np.int64)
L = sb.lib()
L.sb2_engine_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert L.sb2_engine_read_trace(eng._h, out.ctypes.data, 64) == 0
names = ["I:bar_h woke", "I:r,z issued", "I:bar_rh woke", "I:c issued", "G:bar_r woke", "G:ld r done", "G:rh stored",
         "G:rh arrived", "G:bar_z woke", "G:z done", "G:bar_c woke", "G:h stored", "G:h arrived"]
t = out.reshape(4, 16)
for s in range(1):
    base = t[s, 0]
    ev = sorted([(t[s, i] - base, names[i]) for i in range(13)])
    print("step %d (next step starts at +%d)" % (100 + s, t[s + 1, 0] - base))
    for dt, nm in ev:
        print("   +%5d  %s" % (dt, nm))
