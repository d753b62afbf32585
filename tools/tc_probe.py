#!/usr/bin/env python
"""GPU probe for the tcgen05 path: one UMMA tile product against the host (validates the
shared-memory / instruction descriptors and operand layout) and the round-trip latency of
issue -> complete -> commit -> mbarrier wake for the scan's per-step MMA batch."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb

L = sb.lib()
L.sb2_tc_selftest.restype = C.c_int
L.sb2_tc_selftest.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
cases = [(int(a.split("x")[0]), int(a.split("x")[1])) for a in sys.argv[1:]] or [(96, 16), (96, 32), (112, 16), (96, 64)]
for K, N in cases:
    err = C.c_float(-1)
    cyc = (C.c_longlong * 3)()
    reps = 200
    rc = L.sb2_tc_selftest(K, N, reps, C.byref(err), cyc)
    print("selftest K=%d N=%d rc=%d max_abs_err=%.3e  cycles/rep=%.1f issue/rep=%.1f ld=%d  (%s)"
          % (K, N, rc, err.value, cyc[0] / reps, cyc[1] / reps, cyc[2], sb.last_error()), flush=True)
