"""CPU tests of the product's host side through the C-ABI (no GPU calls): exported symbols,
containers, signal preparation, registry, integer post-processing, convolution plan, and the
loud failure when no CUDA device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, bundled_signal
from oracle.oracle import synthetic_read


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "scrappie_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*(?:\([^()]*\)[^;{}()]*)*\)\s*;", hdr)
    return sorted(set(n for n in names if n not in ("defined",) and not n.endswith("_ptr")))


def test_library_exports_every_declared_symbol(sb):
    L = C.CDLL(sb.LIB_PATH)
    names = declared_functions()
    assert len(names) > 45
    for must in ("nanonet_rgrgr_r94_posterior", "nanonet_rnnrf_r94_transitions", "decode_transducer", "decode_crf",
                 "overlapper", "crfpath_to_basecall", "homopolymer_path", "trim_and_segment_raw",
                 "medmad_normalise_array", "mat_from_array", "free_scrappie_matrix", "get_raw_model_stride_from_string",
                 "sb2_basecall_batch", "sb2_batch_forward", "sb2_batch_decode", "nanonet_posterior", "nanonet_raw_posterior",
                 "posterior_crf", "map_to_sequence_viterbi", "map_to_sequence_forward_banded", "encode_bases_to_integers",
                 "sb2_basecall_raw_batch", "sb2_prepare_reads", "sb2_events_posterior_batch", "sb2_multi_stream_time"):
        assert must in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_matrix_container(sb):
    L = sb.lib()
    m = L.make_scrappie_matrix(1025, 7)
    c = m.contents
    assert (c.nr, c.nrq, c.nc, c.stride) == (1025, 257, 7, 1028)
    assert C.addressof(c.f.contents) % 16 == 0
    a = np.ctypeslib.as_array(c.f, shape=(7, 1028))
    assert not a.any()
    assert not L.free_scrappie_matrix(m)
    x = np.arange(15, dtype=np.float32)
    m = L.mat_from_array(x.ctypes.data_as(C.POINTER(C.c_float)), 5, 3)
    a = np.ctypeslib.as_array(m.contents.f, shape=(3, 8))
    assert np.array_equal(a[:, :5].ravel(), x) and not a[:, 5:].any()
    L.free_scrappie_matrix(m)
    assert not L.make_scrappie_matrix(0, 3)


def test_registry(sb):
    L = sb.lib()
    for name, enum, stride in (("raw_r94", 0, 5), ("rgrgr_r94", 1, 5), ("rgrgr_r941", 2, 5), ("rgrgr_r10", 3, 5), ("rnnrf_r94", 4, 1)):
        assert L.get_raw_model(name.encode()) == enum
        assert L.raw_model_string(enum).decode() == name
        assert L.get_raw_model_stride(enum) == stride
        assert sb.get_model_stride(name) == stride
    assert L.get_raw_model(b"nonsense") == 5
    assert L.get_raw_model_stride_from_string(b"nonsense") == -1
    with pytest.raises(ValueError):
        sb.get_model_stride("nonsense")
    assert L.get_homopolymer_calculation(b"mean") == 1 and L.get_homopolymer_calculation(b"nochange") == 0
    assert L.get_homopolymer_calculation(b"x") == 2
    assert L.get_posterior_function(1)


def test_signal_prep_upstream_vectors(sb, golden):
    """Same assertions as src/test/test_scrappie_signal.c:59-103, on the product's host code."""
    g = golden.upstream_signal
    raw = ((g["raw"] + g["offset"]) * (g["range"] / g["digitisation"])).astype(np.float32)
    rt = sb.RawTable(raw).trim(200, 10, 100, 0.0)
    np.testing.assert_allclose(rt.data(as_numpy=True), g["trimmed"], atol=1e-4, rtol=0)
    rt.scale()
    np.testing.assert_allclose(rt.data(as_numpy=True), g["normalised"], atol=1e-5, rtol=0)


def test_signal_prep_bit_exact_vs_oracle(sb, oracle, golden):
    for i in range(3):
        raw = bundled_signal(golden, i)
        rt = sb.RawTable(raw).trim()
        assert (rt.start, rt.end) == oracle.trim_and_segment(raw)
        assert [rt.start, rt.end] == list(golden.ref_reads["r%d_trim" % i])
        rt.scale()
        assert np.array_equal(rt.data(as_numpy=True), oracle.medmad_normalise(raw[rt.start:rt.end]))
    rng = np.random.default_rng(3)
    L = sb.lib()
    fp = C.POINTER(C.c_float)
    for n in (1, 2, 3, 10, 11, 100, 101):
        x = rng.normal(size=n).astype(np.float32)
        assert L.medianf(x.ctypes.data_as(fp), n) == oracle.lib.sb2o_medianf(x.ctypes.data_as(fp), n)
        assert L.madf(x.ctypes.data_as(fp), n, None) == oracle.lib.sb2o_madf(x.ctypes.data_as(fp), n, None)
    assert L.medianf(np.array([3, 1, 2], dtype=np.float32).ctypes.data_as(fp), 3) == 2.0       # src/test/test_util.c
    assert L.medianf(np.array([4, 1, 2, 3], dtype=np.float32).ctypes.data_as(fp), 4) == 2.5


def test_trim_empty_read_fails(sb):
    with pytest.raises(RuntimeError):
        sb.RawTable(np.ones(150, dtype=np.float32)).trim()


def test_overlapper_and_crfpath(sb, oracle, golden):
    L = sb.lib()
    ip = C.POINTER(C.c_int)
    g = golden.ref_reads
    for i in range(3):
        path = np.ascontiguousarray(g["r%d_rgrgr_r94_path" % i])
        pos = np.zeros(path.size, dtype=np.int32)
        call = sb._take_string(L.overlapper(path.ctypes.data_as(ip), path.size, 1024, pos.ctypes.data_as(ip)))
        assert call == str(g["r%d_rgrgr_r94_bases" % i])
        ocall, opos = oracle.overlapper(path, 1024)
        assert call == ocall and np.array_equal(pos, opos)
        cpath = np.ascontiguousarray(g["r%d_rnnrf_r94_path" % i])
        call = sb._take_string(L.crfpath_to_basecall(cpath.ctypes.data_as(ip), cpath.size - 1, pos.ctypes.data_as(ip)))
        assert call == str(g["r%d_rnnrf_r94_bases" % i])
    stays = np.full(10, -1, dtype=np.int32)
    assert not L.overlapper(stays.ctypes.data_as(ip), 10, 1024, None)      # all stays -> NULL
    assert not L.overlapper(None, 10, 1024, None)


def test_homopolymer_path_vs_oracle(sb, oracle, golden):
    """homopolymer_path needs the posterior: use the full synthetic fixtures and a random one."""
    L = sb.lib()
    ip = C.POINTER(C.c_int)
    syn = golden.ref_synthetic
    for key in ("rgrgr_r94_1000", "rgrgr_r94_1003", "rgrgr_r94_997"):
        post = syn[key + "_post"]
        _, vit = oracle.decode_transducer(post, 1025)
        want = oracle.homopolymer_path(post, 1025, vit)
        m = sb.ScrappyMatrix.from_numpy(post, 1025)
        got = vit.copy()
        assert L.homopolymer_path(m.data(), got.ctypes.data_as(ip), 1) == 0
        assert np.array_equal(got, want)
        assert np.array_equal(want, syn[key + "_path"])
        same = vit.copy()
        L.homopolymer_path(m.data(), same.ctypes.data_as(ip), 0)
        assert np.array_equal(same, vit)
    # synthetic path full of homopolymer runs
    rng = np.random.default_rng(11)
    post = np.log(rng.dirichlet(np.ones(1025) * 0.05, size=400).astype(np.float32) + 1e-6).astype(np.float32)
    post = np.hstack([post, np.zeros((400, 3), dtype=np.float32)])
    path = np.full(401, -1, dtype=np.int32)
    for start, base in ((5, 0), (60, 1), (130, 2), (200, 3), (300, 0)):
        rep = sum(base * 4 ** k for k in range(5))
        path[start] = (rep * 4 + (base + 1) % 4) % 1024 if False else ((base + 1) % 4) * 256 + rep // 4
        path[start + 1:start + 12:2] = rep
    want = oracle.homopolymer_path(post, 1025, path)
    m = sb.ScrappyMatrix.from_numpy(post, 1025)
    got = path.copy()
    L.homopolymer_path(m.data(), got.ctypes.data_as(ip), 1)
    assert np.array_equal(got, want)


def naive_reference_conv_columns(n, winlen, stride):
    """Which (x0, tap0, ntap) products the reference's convolution() adds into each column --
    enumerated in Python from src/layers.c:190-241 independently of the C planner."""
    padL, padR = (winlen - 1) // 2, winlen // 2
    ncol = -(-n // stride)
    cols = {c: [] for c in range(ncol)}
    for w in range(0, padL, stride):
        cols[w // stride].append((0, padL - w, winlen - (padL - w)))
    ncolL = -(-padL // stride)
    shift = ncolL * stride - padL
    nstepC = -(-winlen // stride)
    nstepX = stride * nstepC
    for w in range(0, winlen, stride):
        for j in range(max(0, (n - shift - w)) // nstepX):
            c = w // stride + ncolL + j * nstepC
            if c < ncol:
                cols[c].append((shift + w + j * nstepX, 0, winlen))
    maxcol, rem = (n - shift) // nstepX, (n - shift) % nstepX
    colR = ncolL + nstepC * (maxcol - 1) + rem // stride + 1
    startR = stride - (padL + n - winlen) % stride - 1
    for w in range(startR, padR, stride):
        c = colR + w // stride
        if 0 <= c < ncol:
            cols[c].append((n - winlen + 1 + w, 0, winlen - 1 - w))
    return ncol, cols


@pytest.mark.parametrize("winlen,stride", [(19, 5), (11, 1), (11, 5), (9, 2), (7, 3)])
def test_conv_plan_matches_reference_indexing(sb, winlen, stride):
    padL = (winlen - 1) // 2
    for n in list(range(4 * winlen, 4 * winlen + 64)) + [997, 1000, 1003, 4000, 4001, 4004, 12345]:
        first, ncol, plan = sb.conv_plan(n, winlen, stride)
        want_ncol, want = naive_reference_conv_columns(n, winlen, stride)
        assert ncol == want_ncol
        for c in range(ncol):
            if c >= first:
                assert sorted(plan[c]) == sorted(want[c]), (n, c)
            else:   # the kernel's arithmetic rule for non-tail columns
                x0, tap0 = c * stride - padL, 0
                if x0 < 0:
                    tap0, x0 = -x0, 0
                assert want[c] == [(x0, tap0, winlen - tap0)], (n, c)


def test_conv_plan_quirk_n_multiple_of_stride(sb):
    """n % 5 == 0 -> out[T-2] gets the window of column T-1, out[T-1] = bias only."""
    first, ncol, plan = sb.conv_plan(4000, 19, 5)
    assert ncol == 800
    assert plan[799] == []
    assert plan[798] == [(4000 - 19 + 1 + 4, 0, 14)]
    first, ncol, plan = sb.conv_plan(4001, 19, 5)
    assert plan[800] == [(4001 - 19 + 1 + 8, 0, 10)]
    with pytest.raises(RuntimeError):
        sb.conv_plan(10, 19, 5)


def test_no_cpu_fallback(sb):
    """Without a GPU every GPU-backed entry point must fail loudly, not compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no usable CUDA device|failed"):
        sb.Engine(0)
    rt = sb.RawTable(synthetic_read(1, 500))
    with pytest.raises(RuntimeError):
        sb.calc_post(rt, "rgrgr_r94")
    m = sb.ScrappyMatrix.from_numpy(np.zeros((10, 1028), dtype=np.float32), 1025)
    path = np.zeros(11, dtype=np.int32)
    score = sb.lib().decode_transducer(m.data(), 0, 0, 2, path.ctypes.data_as(C.POINTER(C.c_int)), False)
    assert score != score       # NAN
    assert "CUDA" in sb.last_error()


def test_product_never_touches_oracle():
    """The product library and package must not reference oracle/ (judge's rule)."""
    pkg = os.path.join(ROOT, "scrappie_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".c", ".cu", ".h", ".cuh")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn)).read()
                assert "sb2o_" not in txt and "liboracle" not in txt and "scrappie_oracle" not in txt, fn
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_event_features_bit_exact_vs_oracle(sb, oracle):
    """nanonet_features_from_events (host; Kahan sums + RSQRTPS, src/nnfeatures.c:47-115) == the oracle restatement
    on this CPU, bit for bit; sub-range tables use et.start / et.end like the reference."""
    from oracle.oracle import synthetic_events
    for n in (2, 3, 50, 1000):
        ev = synthetic_events(100 + n, n)
        got = sb.event_features(ev)
        assert got.shape == (n, 4)
        assert np.array_equal(got.view(np.uint32), oracle.event_features(ev).view(np.uint32))
    ev = synthetic_events(5, 60)
    sub = sb.EventTable(ev, start=10, end=45)
    assert np.array_equal(sb.event_features(sub).view(np.uint32), oracle.event_features(ev[10:45]).view(np.uint32))
    empty = sb.EventTable(ev, start=7, end=7)
    assert not sb.lib().nanonet_features_from_events(empty.table, True)


def test_encode_bases_and_bounds(sb, reference):
    """encode_bases_to_integers (src/scrappie_seq_helpers.c:53-75) and are_bounds_sane (src/decode.c:1638-1691)."""
    seq = "ACGTTGCAAACCGGTTacgt"
    for k in (1, 3, 5):
        got = sb.encode_bases(seq, k)
        want = [sum("ACGT".index(c) * 4 ** (k - 1 - j) for j, c in enumerate(seq.upper()[i:i + k])) for i in range(len(seq) - k + 1)]
        assert got.tolist() == want
        if reference is not None:
            assert np.array_equal(got, reference.encode_bases(seq, k))
    with pytest.raises(RuntimeError):
        sb.encode_bases("ACGNT", 2)
    sp = C.POINTER(C.c_size_t)

    def sane(lo, hi, seqlen):
        lo = np.ascontiguousarray(lo, dtype=np.uintp)
        hi = np.ascontiguousarray(hi, dtype=np.uintp)
        return bool(sb.lib().are_bounds_sane(lo.ctypes.data_as(sp), hi.ctypes.data_as(sp), lo.size, seqlen))
    assert sane([0, 0, 1, 2], [2, 3, 4, 5], 5)
    assert sane([0, 2], [2, 5], 5)                   # touching, not overlapping: a step is still possible
    assert not sane([1, 1], [3, 5], 5)               # first band must start at 0
    assert not sane([0, 1], [3, 4], 5)               # last band must end at seqlen
    assert not sane([0, 4], [3, 5], 5)               # gap between consecutive bands
    assert not sane([0, 2, 1], [3, 4, 5], 5)         # low bounds must not decrease
    assert not sane([0, 1, 2], [4, 3, 5], 5)         # high bounds must not decrease
    assert not sane([0, 1], [6, 5], 5)               # beyond the sequence


def _build_example(sb, tmp_path):
    import subprocess
    exe = str(tmp_path / "raw_basecall")
    libdir = os.path.dirname(sb.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "raw_basecall.c"), "-L" + libdir, "-lscrappie_b200",
                    "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    return exe


def test_c_caller_compiles_and_runs_host_side(sb, tmp_path):
    """The header is plain C99 and a C program links against the library: examples/raw_basecall.c in its host-only
    mode (registry, containers, trim + scale) -- no GPU call."""
    import subprocess
    exe = _build_example(sb, tmp_path)
    out = subprocess.run([exe, "rnnrf_r94"], check=True, capture_output=True, text=True).stdout
    assert "model rnnrf_r94 stride 1" in out and "trimmed to [200, 390)" in out
    assert subprocess.run([exe, "no_such_model"], capture_output=True).returncode != 0


def test_detect_events_bit_exact(sb, golden, reference):
    """detect_events (host; src/event_detection.c:24-320) against the compiled reference's event tables: fixtures for
    synthetic signals and the bundled reads, the live reference (where built) for odd lengths and parameters."""
    import hashlib
    g = golden.ref_detect
    for seed, n in ((1, 4000), (2, 1234)):
        x = (synthetic_read(seed, n) * 12 + 90).astype(np.float32)
        got = sb.detect_events(x).astype(np.float32)
        assert np.array_equal(got.view(np.uint32), g["syn_%d_events" % n].view(np.uint32))
    for i in range(3):
        got = sb.detect_events(bundled_signal(golden, i)).astype(np.float32)
        assert got.shape[0] == int(g["r%d_nevent" % i])
        assert hashlib.md5(np.ascontiguousarray(got).tobytes()).hexdigest() == str(g["r%d_md5" % i])
        assert got[0, 0] == 0 and np.all(np.diff(got[:, 0]) == got[:-1, 1])            # contiguous events from 0
    if reference is not None:
        for n, kw in ((50, {}), (13, {}), (5000, dict(w1=4, w2=9, t1=2.0, t2=5.0, peak_height=0.5)),
                      (5000, dict(w1=2, w2=3, t1=0.5, t2=1.0, peak_height=0.05))):
            x = (synthetic_read(90 + n, n) * 12 + 90).astype(np.float32)
            par = sb.DetectorParam(kw.get("w1", 3), kw.get("w2", 6), kw.get("t1", 1.4), kw.get("t2", 9.0), kw.get("peak_height", 0.2))
            got = sb.detect_events(x, par).astype(np.float32)
            want = reference.detect_events(x, **kw).astype(np.float32)
            assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    flat = np.full(300, 90.0, dtype=np.float32)                   # no boundary at all: one event over the whole signal
    ev = sb.detect_events(flat)
    assert ev.shape == (1, 4) and ev[0, 0] == 0 and ev[0, 1] == 300 and ev[0, 2] == 90.0 and ev[0, 3] == 0.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    timed on the compiled reference where it is built (oracle/_ref), else on the oracle port.  Tiny sample here."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "8",
                          "--samples", "1000", "--steps", "1", "--warmup", "1"], check=True, capture_output=True, text=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["metric"] == "raw samples/sec (rgrgr_r94)" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(line["e2e"]["value"] - line["value"]) < 1e-6 * line["value"]


def test_bench_resident_plan():
    """bench.py's plan of resident workspaces: whatever the workload, every batch of a step is computed exactly `steps`
    times in the timed region, the sharded shard is streamed through a bounded number of workspaces (config 5 at N = 2
    is 196 batches per rank: they do not fit in HBM at once), and the mixed-length workload gives its long batches more
    copies than its short ones within the same object budget."""
    import collections
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    def runs(obj_group, reps):
        c = collections.Counter()
        for k, r in zip(obj_group, reps):
            c[k] += r
        return c

    # fixed: 4 batches, 4 buffer sets, 20 steps
    og, reps, mult, nsets, steps, warmup = bench.plan_resident([[4000] * 256] * 4, "fixed", 20, 5, 4, 256)
    assert len(og) == 16 and mult == [1] * 4 and nsets == 4 and steps == 20
    assert all(v == 20 for v in runs(og, reps).values()) and sorted(set(og)) == [0, 1, 2, 3] and og[:4] == [0, 1, 2, 3]
    # fewer steps than sets
    og, reps, mult, nsets, steps, _ = bench.plan_resident([[4000] * 256] * 4, "fixed", 2, 1, 4, 256)
    assert nsets == 2 and all(v == 2 for v in runs(og, reps).values())
    # sharded: 50 000 reads = 195 full batches + one of 80 reads
    groups = [[4000] * 256] * 195 + [[4000] * 80]
    og, reps, mult, nsets, steps, warmup = bench.plan_resident(groups, "sharded", 2, 1, 1, 256)
    assert len(og) == 17 and nsets == 1 and steps == 2 and sum(mult) == 196 and mult[-1] == 1 and og[-1] == 195
    assert reps == [m * steps for m in mult]
    og, reps, mult, _, steps, _ = bench.plan_resident([[4000] * 100], "sharded", 2, 1, 1, 256)     # less than one batch
    assert og == [0] and mult == [1] and reps == [steps]
    # mixed: the long batch gets the most copies, every batch still runs `steps` times, object budget respected
    groups = [[130000] * 10, [40000] * 26, [9000] * 120, [4900] * 256, [1700] * 66]
    og, reps, mult, nsets, steps, _ = bench.plan_resident(groups, "mixed", 12, 6, 6, 256)
    r = runs(og, reps)
    assert all(r[k] == 12 for k in range(5)) and og[:5] == [0, 1, 2, 3, 4]
    ncopy = collections.Counter(og)
    assert ncopy[0] == 12 and ncopy[0] >= ncopy[1] >= ncopy[2] >= ncopy[3] >= ncopy[4] >= 1
    assert len(og) <= 6 * 5 + 5
