#!/usr/bin/env python
"""Timeline of the concurrent batches of the bench workload (stage boundaries from CUDA events).

    python tools/exp_timeline.py [--reads 1024] [--batch 256] [--nrep 6]
Prints ms/step and, for the last repetition, when each stage of each batch started.  Environment knobs
(SCRAPPIE_B200_DECODE=..., SCRAPPIE_B200_SCAN=...) select kernel generations; run once per setting.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb
from scrappie_b200.synthetic import synthetic_read


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--samples", type=int, default=4000)
    ap.add_argument("--nrep", type=int, default=6)
    ap.add_argument("--model", default="rgrgr_r94")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    eng = sb.Engine(0)
    base = [synthetic_read(1000 + i, a.samples) for i in range(min(a.reads, 64))]
    sigs = [base[i % len(base)] for i in range(a.reads)]
    nb = (a.reads + a.batch - 1) // a.batch
    groups = [sigs[i * a.batch:(i + 1) * a.batch] for i in range(nb)]
    batches = [eng.batch(a.model, [len(s) for s in g]) for g in groups]
    for b, g in zip(batches, groups):
        b.upload(g)
    p = sb.default_params()
    sb.multi_time(batches, p, nrep=3, flush_l2=True)
    ms = sb.multi_time(batches, p, nrep=a.nrep, flush_l2=True)
    tag = " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("SCRAPPIE_B200_"))
    print("[%s] reads %d batch %d: ms/step graph-replayed %.3f (min %.3f), last (eager, staged) %.3f" %
          (tag, a.reads, a.batch, float(np.mean(ms[:-1])), float(np.min(ms[:-1])), float(ms[-1])))
    for nrep in (10, 20):
        tot = sb.multi_stream_time(batches, p, nrep=nrep)
        print("    streaming, %d steps back to back: %.3f ms/step" % (nrep, tot / nrep))
    if not a.quiet:
        names = list(sb.Batch.STAGES) + ["end"]
        print("%-12s" % "stage" + "".join("  batch%-2d" % k for k in range(nb)))
        offs = [b.stage_offsets() for b in batches]
        for i, n in enumerate(names):
            print("%-12s" % n + "".join("  %7.3f" % o[i] for o in offs))


if __name__ == "__main__":
    main()
