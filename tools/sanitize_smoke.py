#!/usr/bin/env python
"""A small pass over every kernel family for compute-sanitizer (memcheck / initcheck / synccheck are 10-100x slower than
a plain run, so the shapes are tiny): rgrgr_r94 and rnnrf_r94 basecalls of a few reads (small-batch scan, RPG = 4), a
64-read batch of short reads (RPG = 8 scan, affine / head tcgen05 kernels, warp decoder, device finishing), the
raw-signal entry point (trim + med-MAD kernels), posterior_crf, the events model and map_to_sequence."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scrappie_b200 as sb  # noqa: E402
from scrappie_b200.synthetic import synthetic_read  # noqa: E402

eng = sb.Engine(0)
few = [synthetic_read(10 + i, 300 + 7 * i) for i in range(3)]
for model in ("rgrgr_r94", "rnnrf_r94", "raw_r94"):
    calls = eng.basecall_batch(model, few)
    assert all(c[0] for c in calls), model
many = [synthetic_read(100 + i, 200 + (i % 5)) for i in range(64)]
for model in ("rgrgr_r94", "rnnrf_r94"):
    calls = eng.basecall_batch(model, many)
    assert all(c[0] for c in calls), model
    calls2 = eng.basecall_batch(model, many)            # pooled workspace, second use
    assert calls == calls2
longish = [synthetic_read(300 + i, n) for i, n in enumerate((6000, 5203))]   # > 940 blocks: read finishing with the paths in global memory
calls = eng.basecall_batch("rgrgr_r94", longish)
assert all(c[0] for c in calls)
raws = [(synthetic_read(50 + i, n) * np.float32(10) + np.float32(90)).astype(np.float32) for i, n in enumerate((1500, 1300))]
res = eng.basecall_raw_batch("rgrgr_r94", raws)
assert all(r[0] for r in res)
post = sb.calc_post(sb.RawTable(few[0]), "rnnrf_r94")
assert sb.posterior_crf(post).shape[1] == 5
post = sb.calc_post(sb.RawTable(few[1]), "rgrgr_r94")
seq, score, pos = sb.decode_post(post, "rgrgr_r94")
assert seq
score, _ = sb.map_post_to_sequence(post, seq, viterbi=True)
assert score == score
print("sanitize smoke ok: %d launches" % eng.launches)
eng.close()
