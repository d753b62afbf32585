#!/usr/bin/env python
"""Diagnose a base-sequence mismatch of bench.py's parity gate: rebuilds rank 0's shard of the config-5 workload at
N = 8, finds the mismatching read and says whether the GPU's posterior is within tolerance of the reference's and
whether each side's decoder, given the OTHER side's posterior, reproduces the other side's bases (a near-tie in the
Viterbi recursion resolved differently by posteriors 1e-5 apart) -- or whether something else is wrong."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scrappie_b200 as sb  # noqa: E402
from scrappie_b200.sharding import shard_reads  # noqa: E402
from scrappie_b200.synthetic import synthetic_read  # noqa: E402
from oracle.oracle import Oracle, Reference  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
owned = shard_reads([4000] * 100000, 0, world)
distinct = [synthetic_read(1000 + i, 4000) for i in range(2048)]
sigs = [distinct[int(i) % 2048] for i in owned]
groups = [sigs[i:i + 256] for i in range(0, len(sigs), 256)]
rng = np.random.default_rng(17)
flat = [(k, r) for k in range(len(groups)) for r in range(len(groups[k]))]
pick = [flat[i] for i in sorted(rng.choice(len(flat), size=min(48, len(flat)), replace=False))]
ref, orc = Reference(), Oracle()
eng = sb.Engine(0)
seen = {}
for n, (k, r) in enumerate(pick):
    sig = groups[k][r]
    want = ref.basecall_raw("rgrgr_r94", sig)[2]
    got_batch = seen.setdefault(k, eng.basecall_batch("rgrgr_r94", groups[k]))[r][0]
    if got_batch == want:
        continue
    seed = 1000 + int(owned[k * 256 + r]) % 2048
    print("pick %d: batch %d read %d (seed %d): GPU-in-batch != reference; len %d vs %d" % (n, k, r, seed, len(got_batch), len(want)))
    alone = eng.basecall_batch("rgrgr_r94", [sig])[0][0]
    print("  GPU alone == GPU in batch:", alone == got_batch, "| GPU alone == reference:", alone == want)
    print("  oracle port == reference:", orc.basecall_raw("rgrgr_r94", sig)[2] == want)
    b = eng.batch("rgrgr_r94", [len(sig)])
    b.upload([sig]); b.forward(); b.decode()
    gpost = b.posterior(0)
    rpost = ref.posterior("rgrgr_r94", sig)
    opost = orc.posterior("rgrgr_r94", sig)
    print("  max |GPU - reference| posterior (log space) %.3g; |oracle - reference| %.3g; |GPU - oracle| %.3g" % (
        np.abs(gpost[:, :1025] - rpost[:, :1025]).max(), np.abs(opost[:, :1025] - rpost[:, :1025]).max(), np.abs(gpost[:, :1025] - opost[:, :1025]).max()))
    gpath = b.paths()[0][0]
    s_ref_on_g, p_ref_on_g = ref.decode_transducer(gpost, 1025)          # the reference's decoder on the GPU's posterior
    print("  reference decoder on the GPU posterior == GPU path:", np.array_equal(p_ref_on_g, gpath))
    s_r, p_r = ref.decode_transducer(rpost, 1025)
    diff = np.flatnonzero(p_r != gpath)
    print("  paths differ at %d of %d blocks, first %s; Viterbi scores GPU %.6f reference %.6f" % (diff.size, len(gpath), diff[:6], b.paths()[1][0], s_r))
    # first differing base neighbourhood
    m = next((i for i, (x, y) in enumerate(zip(alone, want)) if x != y), min(len(alone), len(want)))
    print("  bases differ from position %d: GPU ...%s... reference ...%s..." % (m, alone[max(0, m - 8):m + 12], want[max(0, m - 8):m + 12]))
    b.close()
print("checked %d picks" % len(pick))
