#!/usr/bin/env python
"""Print posterior / per-layer error of the CUDA path against the oracle and (when present)
the compiled reference, for the scan implementation selected by $SCRAPPIE_B200_SCAN."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scrappie_b200 as sb
from scrappie_b200.synthetic import synthetic_read
from oracle.oracle import Oracle, Reference, reference_available

o = Oracle()
ref = Reference() if reference_available() else None
eng = sb.Engine(0)
mode = os.environ.get("SCRAPPIE_B200_SCAN", "ffma")
for model, lens in (("rgrgr_r94", [4000, 4000, 2503, 1003]), ("rnnrf_r94", [1200, 403])):
    sigs = [synthetic_read(300 + i, n) for i, n in enumerate(lens)]
    b = eng.batch(model, lens)
    b.keep_layers()
    b.upload(sigs)
    b.forward()
    b.decode()
    paths, scores = b.paths()
    H = 96 if model != "rnnrf_r94" else 112
    ns = b.nstate
    for i, s in enumerate(sigs):
        post = b.posterior(i)
        opost, layers = o.posterior(model, s, layers=True)
        lay_err = [float(np.abs(b.layer(l, i, H) - layers[l]).max()) for l in range(6)]
        e_log = float(np.abs(post[:, :ns] - opost[:, :ns]).max())
        line = "%s %-9s n=%-5d log-err vs oracle %.2e" % (mode, model, len(s), e_log)
        if model != "rnnrf_r94":
            e_prob = float(np.abs(np.exp(post[:, :ns]) - np.exp(opost[:, :ns])).max())
            line += "  prob-err %.2e" % e_prob
        if ref is not None:
            rpost = ref.posterior(model, s)
            line += "  | vs reference %.2e (oracle vs reference %.2e)" % (
                float(np.abs(post[:, :ns] - rpost[:, :ns]).max()), float(np.abs(opost[:, :ns] - rpost[:, :ns]).max()))
            rs, rp, rb, _ = ref.basecall_raw(model, s, homopolymer=False)
            line += "  path==ref %s" % bool(np.array_equal(rp, paths[i]))
        print(line)
        print("      layer errs: " + " ".join("%.1e" % e for e in lay_err), flush=True)
    b.close()
