"""Deterministic synthetic raw-signal workloads (numpy only).

SURVEY.md section 8(d): piecewise-constant levels ~U(-1.5, 1.5) held for ~Geometric(mean 9)
samples plus N(0, 0.1^2) noise, then med-MAD normalised; read i of a workload uses seed
1000 + i.  Used by bench.py, the tests and the oracle tooling so that every arm sees the
same inputs."""
import numpy as np


def synthetic_read(seed, n=4000):
    rng = np.random.default_rng(seed)
    out = np.empty(n, dtype=np.float32)
    i = 0
    while i < n:
        d = int(rng.geometric(1.0 / 9.0))
        out[i:i + d] = rng.uniform(-1.5, 1.5)
        i += d
    x = out + rng.normal(0.0, 0.1, n).astype(np.float32)
    med = np.median(x)
    mad = np.median(np.abs(x - med)) * np.float32(1.4826)
    return ((x - med) / mad).astype(np.float32)


def lognormal_lengths(nread, seed=7, median=8000.0, sigma=1.0, lo=1000, hi=200000):
    """BASELINE config 4: read lengths ~ LogNormal(log median, sigma) clipped to [lo, hi]."""
    rng = np.random.default_rng(seed)
    return np.clip(rng.lognormal(np.log(median), sigma, nread), lo, hi).astype(np.int64)
