#!/usr/bin/env python
"""Generate tests/golden/ fixtures.  Runs in the BUILD container only (needs
/root/reference and oracle/_ref/libscrappie_ref.so); the fixtures are committed so
the GPU box, which has neither, can check against them.

Sources of truth:
  upstream_*.npz  the reference's own unit-test vectors (src/test/*.crp, text hex floats),
                  re-encoded losslessly as float32/int32 arrays, with the expected values of
                  src/test/test_scrappie_decoding.c:69-98 and test_scrappie_signal.c:59-103.
  reads.npz       int16 Signal + scaling attributes of the three bundled reads/*.fast5
                  (extracted by tools/fast5_min.py; the reference reads them via libhdf5,
                  src/fast5_interface.c:130-217).
  ref_*.npz       outputs of the reference itself (oracle/_ref), 1 BLAS thread, OpenBLAS
                  0.3.31 (scipy bundled), for seeded synthetic reads and the bundled reads.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.oracle import Reference, synthetic_read  # noqa: E402
import fast5_min  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def read_crp(path):
    with open(path) as fh:
        nr, nc = [int(x) for x in fh.readline().split()]
        mat = np.zeros((nc, nr), dtype=np.float32)
        for c in range(nc):
            mat[c] = [float.fromhex(x) for x in fh.readline().split()]
    return mat


def upstream():
    t = os.path.join(REF, "src", "test")
    post = read_crp(os.path.join(t, "posterior_trimmed.crp"))          # [1000, 1025] probabilities
    path = read_crp(os.path.join(t, "path.crp")).astype(np.int32)[:, 0]
    np.savez_compressed(os.path.join(OUT, "upstream_decode.npz"), posterior=post, path=path,
                        score_expected=np.float32(-115.5761), min_prob=np.float32(1e-5))
    raw = read_crp(os.path.join(t, "raw_signal.crp"))[:, 0]
    trimmed = read_crp(os.path.join(t, "trimmed_signal.crp"))[:, 0]
    normalised = read_crp(os.path.join(t, "normalised_signal.crp"))[:, 0]
    np.savez_compressed(os.path.join(OUT, "upstream_signal.npz"), raw=raw, trimmed=trimmed,
                        normalised=normalised, range=np.float32(1373.41), digitisation=np.float32(8192),
                        offset=np.float32(16))


def bundled_reads():
    d = {}
    names = []
    for fn in sorted(os.listdir(os.path.join(REF, "reads"))):
        if not fn.endswith(".fast5"):
            continue
        r = fast5_min.read_raw(os.path.join(REF, "reads", fn))
        key = "r%d" % len(names)
        names.append(fn)
        d[key + "_signal"] = r["signal_i16"]
        d[key + "_meta"] = np.array([r["digitisation"], r["offset"], r["range"]], dtype=np.float64)
        d[key + "_read_id"] = np.array(r["read_id"])
    d["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "reads.npz"), **d)
    return names


def md5(s):
    return hashlib.md5((s + "\n").encode()).hexdigest()


def reference_outputs():
    ref = Reference()
    # --- seeded synthetic reads: full posterior kept (small) -----------------
    syn = {}
    for model, sizes in (("rgrgr_r94", (500, 503, 1000, 1003, 997, 4000)),
                         ("rnnrf_r94", (500, 1003, 4000)),
                         ("rgrgr_r941", (1000,)), ("rgrgr_r10", (600,))):
        for n in sizes:
            x = synthetic_read(1000 + n, n)
            score, path, bases, post = ref.basecall_raw(model, x)
            key = "%s_%d" % (model, n)
            syn[key + "_score"] = np.float32(score)
            syn[key + "_path"] = path
            syn[key + "_bases"] = np.array(bases)
            if n <= 1003 and model != "rgrgr_r10":
                syn[key + "_post"] = post
            else:
                syn[key + "_post_cols"] = np.arange(0, post.shape[0], 37)
                syn[key + "_post_sub"] = post[::37]
            if model in ("rgrgr_r94", "rnnrf_r94") and n <= 1003:
                syn[key + "_conv"] = ref.convolution(model, x)
    np.savez_compressed(os.path.join(OUT, "ref_synthetic.npz"), **syn)

    # --- decoder sweeps on one posterior ------------------------------------
    dec = {}
    post = syn["rgrgr_r94_1003_post"]
    dec["post_key"] = np.array("rgrgr_r94_1003_post")
    for i, pens in enumerate([(0, 0, 2, False), (2, 0, 2, False), (0, 2, 2, False), (0.5, 1.0, 2, True),
                              (0, 0, 100, False), (0, 0, 0.5, True)]):
        s, p = ref.decode_transducer(post, 1025, *pens)
        dec["pens%d" % i] = np.array(pens, dtype=np.float32)
        dec["score%d" % i] = np.float32(s)
        dec["path%d" % i] = p
    np.savez_compressed(os.path.join(OUT, "ref_decode.npz"), **dec)

    # --- bundled reads, CLI defaults (src/scrappie_raw.c:98-121) ---------------
    reads = np.load(os.path.join(OUT, "reads.npz"))
    out = {}
    for i, fn in enumerate(reads["names"]):
        sig = reads["r%d_signal" % i]
        dig, off, rng = [np.float32(v) for v in reads["r%d_meta" % i]]
        raw = ((sig.astype(np.float32) + off) * np.float32(rng / dig)).astype(np.float32)
        se = ref.trim_and_segment(raw)
        assert se is not None
        s, e = se
        x = ref.medmad_normalise(raw[s:e])
        out["r%d_trim" % i] = np.array([s, e])
        out["r%d_norm_md5" % i] = np.array(hashlib.md5(x.tobytes()).hexdigest())
        for model in ("rgrgr_r94", "rnnrf_r94"):
            score, path, bases, post = ref.basecall_raw(model, x)
            k = "r%d_%s" % (i, model)
            out[k + "_score"] = np.float32(score)
            out[k + "_path"] = path
            out[k + "_bases"] = np.array(bases)
            out[k + "_md5"] = np.array(md5(bases))
            cols = np.arange(0, post.shape[0], 257 if model == "rgrgr_r94" else 29)
            out[k + "_post_cols"] = cols
            out[k + "_post_sub"] = post[cols]
            if model == "rgrgr_r94":
                _, p0 = ref.decode_transducer(post, 1025)
                out[k + "_path_nohp"] = p0
            print(fn[-30:], model, "score %.4f len %d md5 %s" % (score, len(bases), md5(bases)))
    np.savez_compressed(os.path.join(OUT, "ref_reads.npz"), **out)


def raw_r94_outputs():
    """nanonet_raw_posterior (interface/scrappie.h:49-51) fixtures: kept in their own file so that adding them
    did not disturb the other fixtures."""
    ref = Reference()
    d = {}
    for n in (500, 1000, 1003):
        x = synthetic_read(1000 + n, n)
        score, path, bases, post = ref.basecall_raw("raw_r94", x)
        key = "raw_r94_%d" % n
        d[key + "_score"] = np.float32(score)
        d[key + "_path"] = path
        d[key + "_bases"] = np.array(bases)
        d[key + "_post_cols"] = np.arange(0, post.shape[0], 5)
        d[key + "_post_sub"] = post[::5]
    np.savez_compressed(os.path.join(OUT, "ref_raw_r94.npz"), **d)


def posterior_crf_outputs():
    """posterior_crf (src/decode.c:928-1012) of the reference on rnnrf_r94 transitions of seeded synthetic reads
    and on hand-made energies with ties / large magnitudes.  The transitions are stored too, so the GPU test does
    not depend on the network's own (tolerance-level) differences."""
    ref = Reference()
    d = {}
    for n in (300, 1501):
        x = synthetic_read(2000 + n, n)
        trans = ref.posterior("rnnrf_r94", x)
        key = "syn_%d" % n
        d[key + "_trans"] = trans
        d[key + "_post"] = ref.posterior_crf(trans)
    rng = np.random.default_rng(77)
    t = np.zeros((64, 28), dtype=np.float32)
    t[:, :25] = rng.normal(0, 4, size=(64, 25)).astype(np.float32)
    t[10:20, :25] = 0.0                  # exact ties
    t[30:34, :25] *= 10.0                # one transition dominates
    d["hand_trans"] = t
    d["hand_post"] = ref.posterior_crf(t)
    np.savez_compressed(os.path.join(OUT, "ref_posterior_crf.npz"), **d)


def events_outputs():
    """nanonet_posterior (interface/scrappie.h:47-48, the events / LSTM model) of the reference on seeded synthetic
    event tables: studentised features (depend on the CPU's RSQRTPS approximation) and subsampled posterior columns."""
    from oracle.oracle import synthetic_events
    ref = Reference()
    d = {}
    for n in (2, 50, 333, 1200):
        ev = synthetic_events(n, n)
        post = ref.events_posterior(ev)
        key = "ev_%d" % n
        d[key + "_events"] = ev
        d[key + "_features"] = ref.event_features(ev)
        d[key + "_post_cols"] = np.arange(0, post.shape[0], 5)
        d[key + "_post_sub"] = post[::5]
    np.savez_compressed(os.path.join(OUT, "ref_events.npz"), **d)


def map_outputs():
    """map_to_sequence_{viterbi,forward}[_banded] of the reference on the posterior of a seeded synthetic read,
    against its own basecall and a perturbed copy; bands = a diagonal band of half-width 30 positions."""
    ref = Reference()
    x = synthetic_read(4242, 3000)
    score, path, bases, post = ref.basecall_raw("rgrgr_r94", x)
    seq = ref.encode_bases(bases, 5)
    seq2 = seq.copy()
    seq2[10] = (seq2[10] + 37) % 1024
    seq2 = np.delete(seq2, [50, 51, 52])
    nb = post.shape[0]
    d = {"post": post, "bases": np.array(bases), "seq": seq, "seq2": seq2}
    for name, sq in (("a", seq), ("b", seq2)):
        grad = sq.size / nb
        hb = 30 * grad
        lo = np.array([max(0, i * grad - hb) for i in range(nb)], dtype=np.uintp)
        hi = np.array([min(sq.size, i * grad + hb) for i in range(nb)], dtype=np.uintp)
        lo[0] = 0
        hi[-1] = sq.size
        d[name + "_low"], d[name + "_high"] = lo, hi
        for pens in ((0.0, 0.0, 4.0), (0.1, 0.3, 2.0)):
            key = "%s_%g_%g_%g" % ((name,) + pens)
            sv, pv = ref.map_to_sequence(post, 1025, sq, *pens, forward=False, want_path=True)
            d[key + "_viterbi"], d[key + "_path"] = np.float32(sv), pv
            d[key + "_forward"] = np.float32(ref.map_to_sequence(post, 1025, sq, *pens, forward=True)[0])
            d[key + "_viterbi_banded"] = np.float32(ref.map_to_sequence(post, 1025, sq, *pens, forward=False, bands=(lo, hi))[0])
            d[key + "_forward_banded"] = np.float32(ref.map_to_sequence(post, 1025, sq, *pens, forward=True, bands=(lo, hi))[0])
    np.savez_compressed(os.path.join(OUT, "ref_map.npz"), **d)


def detect_outputs():
    """detect_events (src/event_detection.c) of the reference: full event tables for two synthetic pA signals,
    count + md5 of the float32 (start, length, mean, stdv) rows for the bundled reads."""
    ref = Reference()
    d = {}
    for seed, n in ((1, 4000), (2, 1234)):
        x = (synthetic_read(seed, n) * 12 + 90).astype(np.float32)
        d["syn_%d_events" % n] = ref.detect_events(x).astype(np.float32)
    reads = np.load(os.path.join(OUT, "reads.npz"))
    for i in range(3):
        sig = reads["r%d_signal" % i]
        dig, off, rng = [np.float32(v) for v in reads["r%d_meta" % i]]
        x = ((sig.astype(np.float32) + off) * np.float32(rng / dig)).astype(np.float32)
        ev = ref.detect_events(x).astype(np.float32)
        d["r%d_nevent" % i] = np.int64(ev.shape[0])
        d["r%d_md5" % i] = np.array(hashlib.md5(np.ascontiguousarray(ev).tobytes()).hexdigest())
    np.savez_compressed(os.path.join(OUT, "ref_detect.npz"), **d)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "detect":
        detect_outputs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "map":
        map_outputs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "events":
        events_outputs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "posterior_crf":
        posterior_crf_outputs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "raw_r94":
        raw_r94_outputs()
        sys.exit(0)
    upstream()
    bundled_reads()
    reference_outputs()
    for fn in sorted(os.listdir(OUT)):
        print("%9d  %s" % (os.path.getsize(os.path.join(OUT, fn)), fn))
