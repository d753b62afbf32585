"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers.

* `Oracle`  -- oracle/liboracle.so, our plain-C restatement (scrappie_oracle.c).
* `Reference` -- oracle/_ref/libscrappie_ref.so, the reference's own sources compiled
  unmodified (present only after `make -C oracle ref` in the build container; the
  built .so travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
WEIGHTS = os.path.join(ROOT, "scrappie_b200", "weights")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


def build(quiet=True):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


class _Tensor(C.Structure):
    _fields_ = [("data", c_float_p), ("nr", C.c_uint32), ("nc", C.c_uint32), ("stride", C.c_uint32)]


class _Model(C.Structure):
    _fields_ = [("conv_stride", C.c_uint32), ("conv_act", C.c_uint32), ("head", C.c_uint32),
                ("residual", C.c_uint32), ("arch", C.c_uint32), ("conv_W", _Tensor), ("conv_b", _Tensor),
                ("iW", _Tensor * 5), ("b", _Tensor * 5), ("sW", _Tensor * 5), ("sW2", _Tensor * 5),
                ("comb_Wf", _Tensor * 2), ("comb_Wb", _Tensor * 2), ("comb_b", _Tensor * 2),
                ("FF_W", _Tensor), ("FF_b", _Tensor)]


class Oracle:
    """Plain-C restatement (oracle/scrappie_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        for f in ("sb2o_expf", "sb2o_logf", "sb2o_logisticf", "sb2o_tanhf", "sb2o_eluf"):
            getattr(L, f).restype = C.c_float
            getattr(L, f).argtypes = [C.c_float]
        L.sb2o_model_from_blob.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Model)]
        L.sb2o_posterior.restype = C.c_size_t
        L.sb2o_posterior.argtypes = [C.POINTER(_Model), c_float_p, C.c_size_t, C.c_float, C.c_float,
                                     C.c_float, C.c_int, c_float_p, C.POINTER(c_float_p)]
        L.sb2o_nstate.restype = C.c_size_t
        L.sb2o_nstate.argtypes = [C.POINTER(_Model)]
        L.sb2o_decode_transducer.restype = C.c_float
        L.sb2o_decode_transducer.argtypes = [c_float_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float,
                                             C.c_float, C.c_float, c_int_p, C.c_int]
        L.sb2o_decode_crf.restype = C.c_float
        L.sb2o_decode_crf.argtypes = [c_float_p, C.c_size_t, C.c_size_t, c_int_p]
        L.sb2o_overlapper.restype = C.c_void_p
        L.sb2o_overlapper.argtypes = [c_int_p, C.c_size_t, C.c_int, c_int_p]
        L.sb2o_posterior_crf.restype = C.c_int
        L.sb2o_posterior_crf.argtypes = [c_float_p, C.c_size_t, C.c_size_t, c_float_p, C.c_size_t]
        L.sb2o_crfpath_to_basecall.restype = C.c_void_p
        L.sb2o_crfpath_to_basecall.argtypes = [c_int_p, C.c_size_t]
        L.sb2o_homopolymer_path.argtypes = [c_float_p, C.c_size_t, C.c_size_t, C.c_size_t, c_int_p]
        L.sb2o_free.argtypes = [C.c_void_p]
        L.sb2o_medianf.restype = C.c_float
        L.sb2o_medianf.argtypes = [c_float_p, C.c_size_t]
        L.sb2o_madf.restype = C.c_float
        L.sb2o_madf.argtypes = [c_float_p, C.c_size_t, c_float_p]
        L.sb2o_medmad_normalise.argtypes = [c_float_p, C.c_size_t]
        L.sb2o_trim_and_segment.argtypes = [c_float_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                            C.c_float, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.sb2o_convolution.argtypes = [c_float_p, C.c_size_t, C.POINTER(_Tensor), C.POINTER(_Tensor),
                                       C.c_size_t, c_float_p]
        L.sb2o_map_to_sequence.restype = C.c_float
        L.sb2o_map_to_sequence.argtypes = [c_float_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_float,
                                           c_int_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), c_int_p]
        L.sb2o_event_features.argtypes = [c_float_p, C.c_size_t, c_float_p]
        L.sb2o_events_posterior.restype = C.c_size_t
        L.sb2o_events_posterior.argtypes = [C.POINTER(_Model), c_float_p, C.c_size_t, C.c_float, C.c_float,
                                            C.c_float, C.c_int, c_float_p]
        self._models = {}

    def map_to_sequence(self, post, nstate, seq, stay_pen=0.0, skip_pen=0.0, local_pen=4.0, forward=False, bands=None,
                        want_path=False):
        """post [nblock, stride] log posterior; seq: k-mer states.  Returns (score, path or None)."""
        post = np.ascontiguousarray(post, dtype=np.float32)
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        nblock, stride = post.shape
        path = np.zeros(nblock, dtype=np.int32) if want_path else None
        lo = hi = None
        if bands is not None:
            lo = np.ascontiguousarray(bands[0], dtype=np.uintp)
            hi = np.ascontiguousarray(bands[1], dtype=np.uintp)
        sp = C.POINTER(C.c_size_t)
        score = self.lib.sb2o_map_to_sequence(_fp(post), nblock, nstate, stride, stay_pen, skip_pen, local_pen, _ip(seq),
                                              seq.size, int(forward), lo.ctypes.data_as(sp) if lo is not None else None,
                                              hi.ctypes.data_as(sp) if hi is not None else None,
                                              _ip(path) if want_path else None)
        return float(score), path

    # -- events model (src/networks.c:146-194) -------------------------------
    def event_features(self, ev):
        """ev: [n, 3] (mean, stdv, length) -> studentised features [n, 4]."""
        ev = np.ascontiguousarray(ev, dtype=np.float32)
        out = np.zeros((ev.shape[0], 4), dtype=np.float32)
        self.lib.sb2o_event_features(_fp(ev), ev.shape[0], _fp(out))
        return out

    def events_posterior(self, ev, min_prob=1e-5, tempW=1.0, tempb=1.0, return_log=True):
        m = self.model("nanonet_events")
        ev = np.ascontiguousarray(ev, dtype=np.float32)
        ns = self.nstate("nanonet_events")
        out = np.zeros((ev.shape[0], 4 * ((ns + 3) // 4)), dtype=np.float32)
        n = self.lib.sb2o_events_posterior(C.byref(m), _fp(ev), ev.shape[0], min_prob, tempW, tempb,
                                           int(return_log), _fp(out))
        assert n == ev.shape[0]
        return out

    # -- models ------------------------------------------------------------
    def model(self, name):
        if name not in self._models:
            blob = np.fromfile(os.path.join(WEIGHTS, name + ".bin"), dtype=np.uint8)
            m = _Model()
            rc = self.lib.sb2o_model_from_blob(blob.ctypes.data, blob.size, C.byref(m))
            assert rc == 0, "bad weight blob for %s" % name
            self._models[name] = (m, blob)
        return self._models[name][0]

    def nstate(self, name):
        return int(self.lib.sb2o_nstate(C.byref(self.model(name))))

    def posterior(self, name, raw, min_prob=1e-5, tempW=1.0, tempb=1.0, return_log=True, layers=False):
        """Returns post[ncol, stride] (and the 6 layer outputs [ncol, H] if layers)."""
        m = self.model(name)
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        ncol = (raw.size + m.conv_stride - 1) // m.conv_stride
        ns = self.nstate(name)
        stride = 4 * ((ns + 3) // 4)
        out = np.zeros((ncol, stride), dtype=np.float32)
        H = m.conv_W.nc
        assert not (layers and m.arch != 0), "per-layer dumps exist only for the rgrgr / rnnrf stack"
        lay = [np.zeros((ncol, H), dtype=np.float32) for _ in range(6)] if layers else None
        lp = (c_float_p * 6)(*[_fp(a) for a in lay]) if layers else None
        n = self.lib.sb2o_posterior(C.byref(m), _fp(raw), raw.size, min_prob, tempW, tempb,
                                    int(return_log), _fp(out), lp)
        assert n == ncol
        return (out, lay) if layers else out

    def convolution(self, name, raw):
        m = self.model(name)
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        ncol = (raw.size + m.conv_stride - 1) // m.conv_stride
        out = np.zeros((ncol, m.conv_W.nc), dtype=np.float32)
        self.lib.sb2o_convolution(_fp(raw), raw.size, C.byref(m.conv_W), C.byref(m.conv_b),
                                  m.conv_stride, _fp(out))
        return out

    # -- decoders ------------------------------------------------------------
    def decode_transducer(self, post, nstate, stay_pen=0.0, skip_pen=0.0, local_pen=2.0, slip=False):
        post = np.ascontiguousarray(post, dtype=np.float32)
        nblock, stride = post.shape
        seq = np.zeros(nblock + 1, dtype=np.int32)
        score = self.lib.sb2o_decode_transducer(_fp(post), nblock, nstate, stride, stay_pen, skip_pen,
                                                local_pen, _ip(seq), int(slip))
        return float(score), seq

    def decode_crf(self, trans):
        trans = np.ascontiguousarray(trans, dtype=np.float32)
        nblock, stride = trans.shape
        path = np.zeros(nblock + 1, dtype=np.int32)
        score = self.lib.sb2o_decode_crf(_fp(trans), nblock, stride, _ip(path))
        return float(score), path

    def posterior_crf(self, trans):
        """(nblock + 1, 8) state probabilities (ACGT-), the reference's matrix layout (stride 8, 5 rows used)."""
        trans = np.ascontiguousarray(trans, dtype=np.float32)
        nblock, stride = trans.shape
        post = np.zeros((nblock + 1, 8), dtype=np.float32)
        rc = self.lib.sb2o_posterior_crf(_fp(trans), nblock, stride, _fp(post), 8)
        assert rc == 0
        return post

    def _take_str(self, ptr):
        if not ptr:
            return None
        s = C.string_at(ptr).decode()
        self.lib.sb2o_free(ptr)
        return s

    def overlapper(self, seq, nkmer):
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        pos = np.zeros(seq.size, dtype=np.int32)
        return self._take_str(self.lib.sb2o_overlapper(_ip(seq), seq.size, nkmer, _ip(pos))), pos

    def crfpath_to_basecall(self, path, npos):
        path = np.ascontiguousarray(path, dtype=np.int32)
        return self._take_str(self.lib.sb2o_crfpath_to_basecall(_ip(path), npos))

    def homopolymer_path(self, post, nstate, path):
        post = np.ascontiguousarray(post, dtype=np.float32)
        path = np.ascontiguousarray(path, dtype=np.int32).copy()
        self.lib.sb2o_homopolymer_path(_fp(post), post.shape[0], nstate, post.shape[1], _ip(path))
        return path

    # -- signal prep -----------------------------------------------------------
    def medmad_normalise(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).copy()
        self.lib.sb2o_medmad_normalise(_fp(x), x.size)
        return x

    def trim_and_segment(self, raw, trim_start=200, trim_end=10, chunk=100, perc=0.0):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        s, e = C.c_size_t(0), C.c_size_t(0)
        rc = self.lib.sb2o_trim_and_segment(_fp(raw), raw.size, trim_start, trim_end, chunk, perc,
                                            C.byref(s), C.byref(e))
        return None if rc else (s.value, e.value)

    def basecall_raw(self, name, raw, min_prob=1e-5, stay_pen=0.0, skip_pen=0.0, local_pen=2.0,
                     slip=False, homopolymer=True):
        """calculate_post equivalent (src/scrappie_raw.c:265-315) on an already trimmed +
        normalised signal.  Returns (score, path, bases, post)."""
        post = self.posterior(name, raw, min_prob=min_prob)
        ns = self.nstate(name)
        if self.model(name).head == 0:
            score, path = self.decode_transducer(post, ns, stay_pen, skip_pen, local_pen, slip)
            if homopolymer:
                path = self.homopolymer_path(post, ns, path)
            bases, _ = self.overlapper(path, ns - 1)
        else:
            score, path = self.decode_crf(post)
            bases = self.crfpath_to_basecall(path, post.shape[0])
        return score, path, bases, post


# ---------------------------------------------------------------------------
# the compiled reference
# ---------------------------------------------------------------------------

class _Mat(C.Structure):
    _fields_ = [("nr", C.c_size_t), ("nrq", C.c_size_t), ("nc", C.c_size_t), ("stride", C.c_size_t),
                ("data", c_float_p)]


class _Event(C.Structure):
    """event_t, src/scrappie_structures.h:8-15."""
    _fields_ = [("start", C.c_uint64), ("length", C.c_float), ("mean", C.c_float), ("stdv", C.c_float),
                ("pos", C.c_int), ("state", C.c_int)]


class _EventTable(C.Structure):
    _fields_ = [("n", C.c_size_t), ("start", C.c_size_t), ("end", C.c_size_t), ("event", C.POINTER(_Event))]


def make_event_table(ev, start=0, end=None):
    """[n, 3] (mean, stdv, length) -> (event_table, keep-alive array)."""
    ev = np.asarray(ev, dtype=np.float32)
    arr = (_Event * ev.shape[0])()
    pos = 0
    for i in range(ev.shape[0]):
        arr[i].start = pos
        arr[i].length = float(ev[i, 2])
        arr[i].mean = float(ev[i, 0])
        arr[i].stdv = float(ev[i, 1])
        arr[i].pos = -1
        arr[i].state = -1
        pos += int(ev[i, 2])
    return _EventTable(ev.shape[0], start, ev.shape[0] if end is None else end, arr), arr


def synthetic_events(seed, n):
    """Event table of a synthetic read: [n, 3] = (mean pA, stdv, length in samples)."""
    rng = np.random.default_rng(seed)
    mean = (90.0 + 12.0 * rng.uniform(-1.5, 1.5, n)).astype(np.float32)
    stdv = (0.3 + np.abs(rng.normal(1.5, 0.5, n))).astype(np.float32)
    length = rng.geometric(1.0 / 9.0, n).astype(np.float32)
    return np.stack([mean, stdv, length], axis=1).astype(np.float32)


class _DetectorParam(C.Structure):
    """detector_param, src/event_detection.h:6-12; defaults :15-21."""
    _fields_ = [("window_length1", C.c_size_t), ("window_length2", C.c_size_t), ("threshold1", C.c_float),
                ("threshold2", C.c_float), ("peak_height", C.c_float)]


class _RawTable(C.Structure):
    _fields_ = [("uuid", C.c_char_p), ("n", C.c_size_t), ("start", C.c_size_t), ("end", C.c_size_t),
                ("raw", c_float_p)]


REF_SO = os.path.join(HERE, "_ref", "libscrappie_ref.so")


def reference_available():
    return os.path.exists(REF_SO)


class Reference:
    """The reference's own C code (oracle/_ref/libscrappie_ref.so), 1 BLAS thread."""

    POSTERIOR = {"raw_r94": "nanonet_raw_posterior", "rgrgr_r94": "nanonet_rgrgr_r94_posterior", "rgrgr_r941": "nanonet_rgrgr_r941_posterior",
                 "rgrgr_r10": "nanonet_rgrgr_r10_posterior", "rnnrf_r94": "nanonet_rnnrf_r94_transitions"}

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        for f in self.POSTERIOR.values():
            fn = getattr(L, f)
            fn.restype = C.POINTER(_Mat)
            fn.argtypes = [_RawTable, C.c_float, C.c_float, C.c_float, C.c_bool]
        L.free_scrappie_matrix.restype = C.c_void_p
        L.free_scrappie_matrix.argtypes = [C.POINTER(_Mat)]
        L.mat_from_array.restype = C.POINTER(_Mat)
        L.mat_from_array.argtypes = [c_float_p, C.c_size_t, C.c_size_t]
        L.decode_transducer.restype = C.c_float
        L.decode_transducer.argtypes = [C.POINTER(_Mat), C.c_float, C.c_float, C.c_float, c_int_p, C.c_bool]
        L.decode_crf.restype = C.c_float
        L.decode_crf.argtypes = [C.POINTER(_Mat), c_int_p]
        L.overlapper.restype = C.c_void_p
        L.overlapper.argtypes = [c_int_p, C.c_size_t, C.c_int, c_int_p]
        L.crfpath_to_basecall.restype = C.c_void_p
        L.crfpath_to_basecall.argtypes = [c_int_p, C.c_size_t, c_int_p]
        L.posterior_crf.restype = C.POINTER(_Mat)
        L.posterior_crf.argtypes = [C.POINTER(_Mat)]
        sp = C.POINTER(C.c_size_t)
        L.map_to_sequence_viterbi.restype = C.c_float
        L.map_to_sequence_viterbi.argtypes = [C.POINTER(_Mat), C.c_float, C.c_float, C.c_float, c_int_p, C.c_size_t, c_int_p]
        L.map_to_sequence_forward.restype = C.c_float
        L.map_to_sequence_forward.argtypes = [C.POINTER(_Mat), C.c_float, C.c_float, C.c_float, c_int_p, C.c_size_t]
        for f in (L.map_to_sequence_viterbi_banded, L.map_to_sequence_forward_banded):
            f.restype = C.c_float
            f.argtypes = [C.POINTER(_Mat), C.c_float, C.c_float, C.c_float, c_int_p, C.c_size_t, sp, sp]
        L.encode_bases_to_integers.restype = C.POINTER(C.c_int)
        L.encode_bases_to_integers.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t]
        L.detect_events.restype = _EventTable
        L.detect_events.argtypes = [_RawTable, _DetectorParam]
        L.nanonet_posterior.restype = C.POINTER(_Mat)
        L.nanonet_posterior.argtypes = [_EventTable, C.c_float, C.c_float, C.c_float, C.c_bool]
        L.nanonet_features_from_events.restype = C.POINTER(_Mat)
        L.nanonet_features_from_events.argtypes = [_EventTable, C.c_bool]
        L.homopolymer_path.argtypes = [C.POINTER(_Mat), c_int_p, C.c_int]
        L.medmad_normalise_array.argtypes = [c_float_p, C.c_size_t]
        L.trim_and_segment_raw.restype = _RawTable
        L.trim_and_segment_raw.argtypes = [_RawTable, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float]
        L.convolution.restype = C.POINTER(_Mat)
        L.convolution.argtypes = [C.POINTER(_Mat), C.POINTER(_Mat), C.POINTER(_Mat), C.c_size_t, C.POINTER(_Mat)]
        blas = C.CDLL(None)
        try:
            L.scipy_openblas_set_num_threads(1)
        except AttributeError:
            pass

    @staticmethod
    def _to_np(mp):
        m = mp.contents
        return np.ctypeslib.as_array(m.data, shape=(m.nc, m.stride)).copy(), int(m.nr)

    def _mat(self, a2d, nr):
        """numpy [nc, stride] -> reference scrappie_matrix with nr rows."""
        a2d = np.ascontiguousarray(a2d, dtype=np.float32)
        packed = np.ascontiguousarray(a2d[:, :nr])
        return self.lib.mat_from_array(_fp(packed), nr, a2d.shape[0])

    def posterior(self, name, raw, min_prob=1e-5, tempW=1.0, tempb=1.0, return_log=True):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        rt = _RawTable(None, raw.size, 0, raw.size, _fp(raw))
        mp = getattr(self.lib, self.POSTERIOR[name])(rt, min_prob, tempW, tempb, return_log)
        assert mp, "reference returned NULL"
        out, nr = self._to_np(mp)
        self.lib.free_scrappie_matrix(mp)
        return out

    def convolution(self, name, raw):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        X = self.lib.mat_from_array(_fp(raw), 1, raw.size)
        W = _Mat.in_dll(self.lib, "_conv_%s_W" % name)
        b = _Mat.in_dll(self.lib, "_conv_%s_b" % name)
        stride = C.c_int.in_dll(self.lib, "conv_%s_stride" % name).value
        mp = self.lib.convolution(X, C.byref(W), C.byref(b), stride, None)
        out, nr = self._to_np(mp)
        self.lib.free_scrappie_matrix(mp)
        self.lib.free_scrappie_matrix(X)
        return out[:, :nr]

    def decode_transducer(self, post, nstate, stay_pen=0.0, skip_pen=0.0, local_pen=2.0, slip=False):
        mp = self._mat(post, nstate)
        seq = np.zeros(post.shape[0] + 1, dtype=np.int32)
        score = self.lib.decode_transducer(mp, stay_pen, skip_pen, local_pen, _ip(seq), slip)
        self.lib.free_scrappie_matrix(mp)
        return float(score), seq

    def decode_crf(self, trans):
        mp = self._mat(trans, 25)
        path = np.zeros(trans.shape[0] + 1, dtype=np.int32)
        score = self.lib.decode_crf(mp, _ip(path))
        self.lib.free_scrappie_matrix(mp)
        return float(score), path

    def encode_bases(self, bases, klen):
        b = bases.encode()
        ptr = self.lib.encode_bases_to_integers(b, len(b), klen)
        out = np.ctypeslib.as_array(ptr, shape=(len(b) - klen + 1,)).copy()
        self.libc.free(ptr)
        return out

    def map_to_sequence(self, post, nstate, seq, stay_pen=0.0, skip_pen=0.0, local_pen=4.0, forward=False, bands=None,
                        want_path=False):
        mp = self._mat(post, nstate)
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        path = None
        if bands is None:
            if forward:
                score = self.lib.map_to_sequence_forward(mp, stay_pen, skip_pen, local_pen, _ip(seq), seq.size)
            else:
                path = np.zeros(post.shape[0], dtype=np.int32) if want_path else None
                score = self.lib.map_to_sequence_viterbi(mp, stay_pen, skip_pen, local_pen, _ip(seq), seq.size,
                                                         _ip(path) if want_path else None)
        else:
            sp = C.POINTER(C.c_size_t)
            lo = np.ascontiguousarray(bands[0], dtype=np.uintp)
            hi = np.ascontiguousarray(bands[1], dtype=np.uintp)
            f = self.lib.map_to_sequence_forward_banded if forward else self.lib.map_to_sequence_viterbi_banded
            score = f(mp, stay_pen, skip_pen, local_pen, _ip(seq), seq.size, lo.ctypes.data_as(sp), hi.ctypes.data_as(sp))
        self.lib.free_scrappie_matrix(mp)
        return float(score), path

    def detect_events(self, raw, w1=3, w2=6, t1=1.4, t2=9.0, peak_height=0.2):
        """[n, 4] array (start, length, mean, stdv) of the reference's event detector."""
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        rt = _RawTable(None, raw.size, 0, raw.size, _fp(raw))
        et = self.lib.detect_events(rt, _DetectorParam(w1, w2, t1, t2, peak_height))
        out = np.array([[et.event[i].start, et.event[i].length, et.event[i].mean, et.event[i].stdv] for i in range(et.n)],
                       dtype=np.float64)
        self.libc.free(C.cast(et.event, C.c_void_p))
        return out

    def events_posterior(self, ev, min_prob=1e-5, tempW=1.0, tempb=1.0, return_log=True):
        et, keep = make_event_table(ev)
        mp = self.lib.nanonet_posterior(et, min_prob, tempW, tempb, return_log)
        assert mp, "reference returned NULL"
        out, nr = self._to_np(mp)
        self.lib.free_scrappie_matrix(mp)
        return out

    def event_features(self, ev):
        et, keep = make_event_table(ev)
        mp = self.lib.nanonet_features_from_events(et, True)
        out, nr = self._to_np(mp)
        self.lib.free_scrappie_matrix(mp)
        return out

    def posterior_crf(self, trans):
        mp = self._mat(trans, 25)
        pp = self.lib.posterior_crf(mp)
        assert pp, "reference returned NULL"
        out, nr = self._to_np(pp)
        self.lib.free_scrappie_matrix(pp)
        self.lib.free_scrappie_matrix(mp)
        return out

    def _take_str(self, ptr):
        if not ptr:
            return None
        s = C.string_at(ptr).decode()
        self.libc.free(ptr)
        return s

    def overlapper(self, seq, nkmer):
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        pos = np.zeros(seq.size, dtype=np.int32)
        return self._take_str(self.lib.overlapper(_ip(seq), seq.size, nkmer, _ip(pos))), pos

    def crfpath_to_basecall(self, path, npos):
        path = np.ascontiguousarray(path, dtype=np.int32)
        pos = np.zeros(path.size, dtype=np.int32)
        return self._take_str(self.lib.crfpath_to_basecall(_ip(path), npos, _ip(pos)))

    def homopolymer_path(self, post, nstate, path):
        mp = self._mat(post, nstate)
        path = np.ascontiguousarray(path, dtype=np.int32).copy()
        self.lib.homopolymer_path(mp, _ip(path), 1)      # HOMOPOLYMER_MEAN (src/homopolymer.h)
        self.lib.free_scrappie_matrix(mp)
        return path

    def medmad_normalise(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).copy()
        self.lib.medmad_normalise_array(_fp(x), x.size)
        return x

    def trim_and_segment(self, raw, trim_start=200, trim_end=10, chunk=100, perc=0.0):
        # the reference frees rt.raw when the range is empty, so hand it a malloc'd copy
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        buf = self.libc.malloc(raw.nbytes)
        C.memmove(buf, raw.ctypes.data, raw.nbytes)
        rt = _RawTable(None, raw.size, 0, raw.size, C.cast(buf, c_float_p))
        out = self.lib.trim_and_segment_raw(rt, trim_start, trim_end, chunk, perc)
        if not out.raw:
            return None
        self.libc.free(buf)
        return int(out.start), int(out.end)

    def basecall_raw(self, name, raw, min_prob=1e-5, stay_pen=0.0, skip_pen=0.0, local_pen=2.0,
                     slip=False, homopolymer=True):
        post = self.posterior(name, raw, min_prob=min_prob)
        if name != "rnnrf_r94":
            ns = {"rgrgr_r10": 4097}.get(name, 1025)
            score, path = self.decode_transducer(post, ns, stay_pen, skip_pen, local_pen, slip)
            if homopolymer:
                path = self.homopolymer_path(post, ns, path)
            bases, _ = self.overlapper(path, ns - 1)
        else:
            score, path = self.decode_crf(post)
            bases = self.crfpath_to_basecall(path, post.shape[0])
        return score, path, bases, post


# the workload generator lives with the product (numpy only); re-exported for the tests
import sys as _sys
if ROOT not in _sys.path:
    _sys.path.insert(0, ROOT)
from scrappie_b200.synthetic import synthetic_read  # noqa: E402,F401
