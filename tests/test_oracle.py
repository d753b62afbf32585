"""CPU tests: the oracle (plain-C restatement) against the reference's own known-answer
vectors (tests/golden/upstream_*.npz, from src/test/*.crp) and against outputs of the
compiled reference (tests/golden/ref_*.npz, made by tools/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import bundled_signal
from oracle.oracle import synthetic_read

LOG_TOL = 1e-4      # posterior, log space


def robustlog(p, min_prob):
    return np.log(np.float32(min_prob) + (np.float32(1) - np.float32(min_prob)) * p).astype(np.float32)


def test_elu_known_answers(oracle):
    # src/test/test_scrappie_elu.c:23-71
    L = oracle.lib
    for x, want in ((-1.0, -0.6321206), (-2.0, -0.8646647), (-3.0, -0.9502129), (-4.0, -0.9816844)):
        assert abs(L.sb2o_eluf(x) - want) < 1e-6
    for x in (0.0, 1.0, 2.5, 100.0):
        assert L.sb2o_eluf(x) == np.float32(x)


def test_exp_log_against_libm(oracle):
    L = oracle.lib
    xs = np.concatenate([np.linspace(-80, 80, 2001), [0.0]]).astype(np.float32)
    for x in xs[::7]:
        want = np.exp(np.float64(x))
        assert abs(L.sb2o_expf(float(x)) - want) <= 5e-7 * want
    # the input clamp of exp_ps (src/sse_mathfun.h:221-222)
    assert L.sb2o_expf(100.0) == L.sb2o_expf(88.3762626647949)
    assert L.sb2o_expf(-100.0) == L.sb2o_expf(-88.3762626647949)
    for x in np.logspace(-30, 30, 400).astype(np.float32):
        assert abs(L.sb2o_logf(float(x)) - np.log(np.float64(x))) <= 2e-7 * max(1.0, abs(np.log(np.float64(x))))
    assert np.isnan(L.sb2o_logf(-1.0))


def test_upstream_decoder_known_answer(oracle, golden):
    """src/test/test_scrappie_decoding.c:69-98: score -115.5761 +-1e-4 and path == path.crp
    shifted by one, with penalties (0, 0, 100) after robustlog(1e-5)."""
    g = golden.upstream_decode
    logpost = np.zeros((1000, 1028), dtype=np.float32)
    logpost[:, :1025] = robustlog(g["posterior"], 1e-5)
    score, path = oracle.decode_transducer(logpost, 1025, 0.0, 0.0, 100.0)
    assert abs(score - float(g["score_expected"])) < 1e-4
    assert np.array_equal(path[1:], g["path"])


def test_upstream_signal_known_answer(oracle, golden):
    """src/test/test_scrappie_signal.c:59-103: trim (MAD segmentation + 200/10) and med-MAD
    normalisation of raw_signal.crp reproduce trimmed_signal.crp / normalised_signal.crp."""
    g = golden.upstream_signal
    raw = ((g["raw"] + g["offset"]) * (g["range"] / g["digitisation"])).astype(np.float32)
    se = oracle.trim_and_segment(raw, 200, 10, 100, 0.0)
    assert se is not None
    s, e = se
    assert e - s == g["trimmed"].size
    np.testing.assert_allclose(raw[s:e], g["trimmed"], atol=1e-4, rtol=0)
    norm = oracle.medmad_normalise(raw[s:e])
    np.testing.assert_allclose(norm, g["normalised"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("key", ["rgrgr_r94_500", "rgrgr_r94_503", "rgrgr_r94_1000", "rgrgr_r94_1003",
                                 "rgrgr_r94_997", "rnnrf_r94_500", "rnnrf_r94_1003", "rgrgr_r941_1000"])
def test_posterior_against_reference_fixture(oracle, golden, key):
    model, n = key.rsplit("_", 1)
    g = golden.ref_synthetic
    x = synthetic_read(1000 + int(n), int(n))
    post = oracle.posterior(model, x)
    ref = g[key + "_post"]
    ns = oracle.nstate(model)
    assert post.shape == ref.shape
    assert np.abs(post[:, :ns] - ref[:, :ns]).max() < LOG_TOL
    score, path, bases, _ = oracle.basecall_raw(model, x)
    assert bases == str(g[key + "_bases"])
    assert np.array_equal(path, g[key + "_path"])
    assert abs(score - float(g[key + "_score"])) < 5e-3


@pytest.mark.parametrize("key", ["rgrgr_r94_500", "rgrgr_r94_503", "rgrgr_r94_1000", "rgrgr_r94_997", "rnnrf_r94_500"])
def test_convolution_edge_behaviour(oracle, golden, key):
    """The stride-5 right-edge quirk (src/layers.c:218-241): n % 5 == 0 moves the last window one
    column early and leaves the final column at bias only; pinned by reference output."""
    model, n = key.rsplit("_", 1)
    conv = oracle.convolution(model, synthetic_read(1000 + int(n), int(n)))
    np.testing.assert_allclose(conv, golden.ref_synthetic[key + "_conv"], atol=2e-6, rtol=0)


def test_rgrgr_r10_subsampled(oracle, golden):
    g = golden.ref_synthetic
    x = synthetic_read(1600, 600)
    score, path, bases, post = oracle.basecall_raw("rgrgr_r10", x)
    assert np.abs(post[g["rgrgr_r10_600_post_cols"]][:, :4097] - g["rgrgr_r10_600_post_sub"][:, :4097]).max() < LOG_TOL
    assert bases == str(g["rgrgr_r10_600_bases"])


@pytest.mark.parametrize("n", [500, 1000, 1003])
def test_raw_r94_against_reference_fixture(oracle, golden, n):
    """Pins the restatement of nanonet_raw_posterior (src/networks.c:196-247)."""
    g = golden.ref_raw_r94
    key = "raw_r94_%d" % n
    x = synthetic_read(1000 + n, n)
    score, path, bases, post = oracle.basecall_raw("raw_r94", x)
    assert np.abs(post[g[key + "_post_cols"]][:, :1025] - g[key + "_post_sub"][:, :1025]).max() < 1e-4
    assert bases == str(g[key + "_bases"])
    assert np.array_equal(path, g[key + "_path"])


def test_decoder_sweeps_exact(oracle, golden):
    g, syn = golden.ref_decode, golden.ref_synthetic
    post = syn[str(g["post_key"])]
    for i in range(6):
        stay, skip, local, slip = [float(v) for v in g["pens%d" % i]]
        score, path = oracle.decode_transducer(post, 1025, stay, skip, local, bool(slip))
        assert score == float(g["score%d" % i])
        assert np.array_equal(path, g["path%d" % i])


def test_crf_decode_exact(oracle, golden):
    syn = golden.ref_synthetic
    post = syn["rnnrf_r94_1003_post"]
    score, path = oracle.decode_crf(post)
    assert np.array_equal(path, syn["rnnrf_r94_1003_path"])
    assert score == float(syn["rnnrf_r94_1003_score"])


def test_bundled_read_end_to_end(oracle, golden):
    """Smallest bundled read (read_ch228_file118) through trim -> normalise -> network -> decode ->
    homopolymer -> overlapper equals the reference's basecall (md5 also listed in SURVEY.md 8c)."""
    g = golden.ref_reads
    i = 2
    raw = bundled_signal(golden, i)
    s, e = oracle.trim_and_segment(raw)
    assert [s, e] == list(g["r%d_trim" % i])
    x = oracle.medmad_normalise(raw[s:e])
    assert hashlib.md5(x.tobytes()).hexdigest() == str(g["r%d_norm_md5" % i])
    score, path, bases, post = oracle.basecall_raw("rgrgr_r94", x)
    assert hashlib.md5((bases + "\n").encode()).hexdigest() == "f0d4357f5e6532d63cef8989647294e3"
    assert bases == str(g["r2_rgrgr_r94_bases"])
    cols = g["r2_rgrgr_r94_post_cols"]
    assert np.abs(post[cols][:, :1025] - g["r2_rgrgr_r94_post_sub"][:, :1025]).max() < LOG_TOL


def test_oracle_matches_live_reference(oracle, reference):
    """Where oracle/_ref is present, compare live on fresh seeds (not only on fixtures)."""
    if reference is None:
        pytest.skip("oracle/_ref not built here")
    for model, n in (("rgrgr_r94", 1234), ("rgrgr_r94", 1235), ("rnnrf_r94", 777)):
        x = synthetic_read(55 + n, n)
        a, b = oracle.posterior(model, x), reference.posterior(model, x)
        ns = oracle.nstate(model)
        assert np.abs(a[:, :ns] - b[:, :ns]).max() < LOG_TOL
        sa, pa, ba, _ = oracle.basecall_raw(model, x)
        sb_, pb, bb, _ = reference.basecall_raw(model, x)
        assert ba == bb and np.array_equal(pa, pb)


@pytest.mark.parametrize("key", ["syn_300", "syn_1501", "hand"])
def test_posterior_crf_exact(oracle, golden, key):
    """posterior_crf restatement vs the compiled reference's output (same libm, same order: bit-exact)."""
    g = golden.ref_posterior_crf
    got = oracle.posterior_crf(g[key + "_trans"])
    assert np.array_equal(got[:, :5], g[key + "_post"][:, :5])


@pytest.mark.parametrize("n", [2, 50, 333, 1200])
def test_events_model_against_reference_fixture(oracle, golden, n):
    """Events (LSTM) model restatement vs the compiled reference (tests/golden/ref_events.npz)."""
    g = golden.ref_events
    ev = g["ev_%d_events" % n]
    feat = oracle.event_features(ev)
    same_cpu = np.array_equal(feat.view(np.uint32), g["ev_%d_features" % n][:, :4].view(np.uint32))
    post = oracle.events_posterior(ev)
    err = np.abs(post[g["ev_%d_post_cols" % n]][:, :1025] - g["ev_%d_post_sub" % n][:, :1025]).max()
    assert err < (1e-4 if same_cpu else 5e-2), (err, same_cpu)


def test_map_to_sequence_exact(oracle, golden):
    """map_to_sequence restatement (all four variants) vs the compiled reference: same libm, same order -> exact."""
    g = golden.ref_map
    post = g["post"]
    for name, sq in (("a", g["seq"]), ("b", g["seq2"])):
        bands = (g[name + "_low"], g[name + "_high"])
        for pens in ((0.0, 0.0, 4.0), (0.1, 0.3, 2.0)):
            key = "%s_%g_%g_%g" % ((name,) + pens)
            sv, pv = oracle.map_to_sequence(post, 1025, sq, *pens, forward=False, want_path=True)
            assert np.float32(sv) == g[key + "_viterbi"] and np.array_equal(pv, g[key + "_path"])
            assert np.float32(oracle.map_to_sequence(post, 1025, sq, *pens, forward=True)[0]) == g[key + "_forward"]
            assert np.float32(oracle.map_to_sequence(post, 1025, sq, *pens, forward=False, bands=bands)[0]) == g[key + "_viterbi_banded"]
            assert np.float32(oracle.map_to_sequence(post, 1025, sq, *pens, forward=True, bands=bands)[0]) == g[key + "_forward_banded"]
