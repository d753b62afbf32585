#!/usr/bin/env python
"""Per-call latency of sb2_basecall_batch on the mixed workload, alone and with N calls in flight (GPU only)."""
import collections
import os
import queue
import sys
import threading
import time

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
order = sorted(range(len(groups)), key=lambda k: -sum(len(s) for s in groups[k]))


def run(nworker, nstep, which=None):
    q = queue.Queue()
    for _ in range(nstep):
        for k in (which or order):
            q.put(k)
    lat = collections.defaultdict(list)

    def worker():
        while True:
            try:
                k = q.get_nowait()
            except queue.Empty:
                return
            t0 = time.perf_counter()
            eng.basecall_prepared("rgrgr_r94", prepared[k], params).close()
            lat[k].append((time.perf_counter() - t0) * 1e3)
    th = [threading.Thread(target=worker) for _ in range(nworker)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return (time.perf_counter() - t0) / nstep, lat


for nworker in (1, 16):
    run(nworker, 2)
    dt, lat = run(nworker, 4)
    print("workers %d: %.1f ms per step; mean latency per batch (reads: ms):" % (nworker, dt * 1e3),
          ", ".join("%d: %.0f" % (len(groups[k]), sum(v) / len(v)) for k, v in sorted(lat.items())))
# only the short-read batches / only the long-read batches in flight
short = [k for k in order if max(len(s) for s in groups[k]) < 20000]
long_ = [k for k in order if k not in short]
for name, which in (("short-read batches only", short), ("long-read batches only", long_)):
    run(16, 2, which)
    dt, lat = run(16, 4, which)
    n = sum(sum(len(s) for s in groups[k]) for k in which)
    print("%s (%d batches, %d samples), 16 workers: %.1f ms per pass = %.3g samples/s;" % (name, len(which), n, dt * 1e3, n / dt),
          ", ".join("%d: %.0f" % (len(groups[k]), sum(v) / len(v)) for k, v in sorted(lat.items())))
