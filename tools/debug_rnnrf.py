#!/usr/bin/env python
"""Where does the rnnrf_r94 scan differ from the oracle?  Per-layer max error and its position for a few reads of a
100-read ragged batch, (GPU only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb  # noqa: E402
from oracle.oracle import Oracle, synthetic_read  # noqa: E402

oracle = Oracle()
rng = np.random.default_rng(11)
lens = [int(x) for x in rng.integers(1000, 1400, size=100)]
lens[9] = 1399
lens[10] = 1000
sigs = [synthetic_read(2000 + i, n) for i, n in enumerate(lens)]
print("lens[:16]", lens[:16])
want = {i: oracle.posterior("rnnrf_r94", sigs[i], layers=True) for i in (0, 4, 5, 6, 9, 10)}
for gen in (6,):
    eng = sb.Engine(0)
    b = eng.batch("rnnrf_r94", lens)
    b.keep_layers()
    b.upload(sigs)
    b.forward()
    for i, (post, layers) in want.items():
        got = b.posterior(i)
        e = np.abs(got[:, :25] - post[:, :25])
        t, k = np.unravel_index(np.argmax(e), e.shape)
        row = ["gen %d read %d T %d: post err %.2e at t=%d row %d |" % (gen, i, lens[i], e.max(), t, k)]
        for l in range(6):
            el = np.abs(b.layer(l, i, 112) - layers[l])
            tl, ul = np.unravel_index(np.argmax(el), el.shape)
            row.append("L%d %.1e@t%d,u%d (|x| %.1f)" % (l, el.max(), tl, ul, np.abs(layers[l]).max()))
        print(" ".join(row))
    b.close()
    eng.close()
