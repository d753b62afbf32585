#!/usr/bin/env python
"""Where does the time of the documented call go on the mixed-length workload (BASELINE config 4)?  GPU only.
Prints, per sb2_basecall_batch call: reads, samples, longest read, wall ms -- first pass (cold pool), second pass (warm),
then the threaded pass bench.py times."""
import os
import queue
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb  # noqa: E402
from scrappie_b200.sharding import lognormal_lengths, plan_batches  # noqa: E402
from scrappie_b200.synthetic import synthetic_read  # noqa: E402

eng = sb.Engine(0)
lens = lognormal_lengths(1024, seed=4)
sigs = [synthetic_read(1000 + i, int(n)) for i, n in enumerate(lens)]
plan = plan_batches(lens, max_reads=256, max_samples=256 * 4096)
groups = [[sigs[i] for i in idx] for idx in plan]
prepared = [eng.prepare_call(g) for g in groups]
params = sb.default_params()
total = sum(len(s) for s in sigs)
print("batches:", [(len(g), sum(len(s) for s in g), max(len(s) for s in g)) for g in groups])
for name in ("cold", "warm", "warm2"):
    t_all = time.perf_counter()
    rows = []
    for k, g in enumerate(groups):
        t0 = time.perf_counter()
        cs = eng.basecall_prepared("rgrgr_r94", prepared[k], params)
        rows.append((len(g), round((time.perf_counter() - t0) * 1e3, 2)))
        cs.close()
    dt = time.perf_counter() - t_all
    print(name, "sequential: %.1f ms total, %.3g samples/s" % (dt * 1e3, total / dt), rows)
order = sorted(range(len(groups)), key=lambda k: -sum(len(s) for s in groups[k]))
for nworker in (8, 16, 24, 32, 48):
    for rep in range(3):
        q = queue.Queue()
        for _ in range(6):
            for k in order:
                q.put(k)

        def worker():
            while True:
                try:
                    k = q.get_nowait()
                except queue.Empty:
                    return
                eng.basecall_prepared("rgrgr_r94", prepared[k], params).close()
        th = [threading.Thread(target=worker) for _ in range(nworker)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = (time.perf_counter() - t0) / 6
        print("threads %d pass %d: %.1f ms per step, %.3g samples/s, reallocs so far %d" % (nworker, rep, dt * 1e3, total / dt, eng.reallocs))
