"""scrappie_b200 -- Python host side of libscrappie_b200.so (ctypes, no torch).

Mirrors the reference's `scrappy` binding for the raw basecalling path
(python/scrappy/__init__.py: RawTable :47-111, ScrappyMatrix :147-192, calc_post :276-299,
decode_post :302-366, basecall_raw :403-430, get_model_stride :390-400) with the same
names, argument meaning and error behaviour, and adds the batch interface
(`Engine`, `Batch`, `basecall_batch`) that the GPU needs.

All numerical work goes through the C-ABI of include/scrappie_b200.h.  There is no
Python or CPU implementation of the network or the decoders in this package: if the
shared library is missing, or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCRAPPIE_B200_LIB selects an alternative build of the library (A/B measurements); the weights stay where they are
# one hardware work queue per batch in flight instead of 8 shared ones (see sb2_engine_create); must be in the environment
# before anything in the process creates a CUDA context
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
LIB_PATH = os.environ.get("SCRAPPIE_B200_LIB") or os.path.join(_HERE, "libscrappie_b200.so")
WEIGHTS_DIR = os.path.join(_HERE, "weights")
if os.environ.get("SCRAPPIE_B200_LIB"):
    os.environ.setdefault("SCRAPPIE_B200_WEIGHTS", WEIGHTS_DIR)     # the library looks for weights next to itself

MODELS = ("raw_r94", "rgrgr_r94", "rgrgr_r941", "rgrgr_r10", "rnnrf_r94")
_MODEL_ENUM = {"raw_r94": 0, "rgrgr_r94": 1, "rgrgr_r941": 2, "rgrgr_r10": 3, "rnnrf_r94": 4}

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


class _Mat(C.Structure):
    _fields_ = [("nr", C.c_size_t), ("nrq", C.c_size_t), ("nc", C.c_size_t), ("stride", C.c_size_t),
                ("f", _f32p)]


class _RawTable(C.Structure):
    _fields_ = [("uuid", C.c_char_p), ("n", C.c_size_t), ("start", C.c_size_t), ("end", C.c_size_t),
                ("raw", _f32p)]


class Params(C.Structure):
    """sb2_params (include/scrappie_b200.h); defaults = `scrappie raw` CLI defaults."""
    _fields_ = [("min_prob", C.c_float), ("tempW", C.c_float), ("tempb", C.c_float),
                ("stay_pen", C.c_float), ("skip_pen", C.c_float), ("local_pen", C.c_float),
                ("allow_slip", C.c_int), ("homopolymer", C.c_int)]


class _Event(C.Structure):
    """event_t (src/scrappie_structures.h:8-15)."""
    _fields_ = [("start", C.c_uint64), ("length", C.c_float), ("mean", C.c_float), ("stdv", C.c_float),
                ("pos", C.c_int), ("state", C.c_int)]


class _EventTable(C.Structure):
    _fields_ = [("n", C.c_size_t), ("start", C.c_size_t), ("end", C.c_size_t), ("event", C.POINTER(_Event))]


class DetectorParam(C.Structure):
    """detector_param (src/event_detection.h:6-21); the defaults are the reference's event_detection_defaults."""
    _fields_ = [("window_length1", C.c_size_t), ("window_length2", C.c_size_t), ("threshold1", C.c_float),
                ("threshold2", C.c_float), ("peak_height", C.c_float)]

    def __init__(self, window_length1=3, window_length2=6, threshold1=1.4, threshold2=9.0, peak_height=0.2):
        super().__init__(window_length1, window_length2, threshold1, threshold2, peak_height)


class EventTable(object):
    """An event_table built from an [n, 3] array of (mean, stdv, length)."""

    def __init__(self, events, start=0, end=None):
        ev = np.asarray(events, dtype=np.float32)
        self._arr = (_Event * ev.shape[0])()
        pos = 0
        for i in range(ev.shape[0]):
            e = self._arr[i]
            e.start, e.length, e.mean, e.stdv, e.pos, e.state = pos, float(ev[i, 2]), float(ev[i, 0]), float(ev[i, 1]), -1, -1
            pos += int(ev[i, 2])
        self.table = _EventTable(ev.shape[0], start, ev.shape[0] if end is None else end, self._arr)


class Trim(C.Structure):
    """sb2_trim: the signal-preparation options of `scrappie raw` (src/scrappie_raw.c:98-121)."""
    _fields_ = [("trim_start", C.c_size_t), ("trim_end", C.c_size_t), ("varseg_chunk", C.c_size_t),
                ("varseg_thresh", C.c_float)]


class _Call(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("score", C.c_float), ("nblock", C.c_size_t), ("nbase", C.c_size_t)]


def _posterior_symbol(model):
    """Exported posterior function of a model (names as in src/networks.h:42-49, interface/scrappie.h:49)."""
    if model == "raw_r94":
        return "nanonet_raw_posterior"
    return "nanonet_%s_%s" % (model, "transitions" if model == "rnnrf_r94" else "posterior")


def build_library(verbose=False):
    """Compile libscrappie_b200.so in-tree with nvcc for sm_100a."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_lib = None


def lib():
    """The loaded shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libscrappie_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    mp = C.POINTER(_Mat)
    sig = {
        "make_scrappie_matrix": (mp, [C.c_size_t, C.c_size_t]),
        "free_scrappie_matrix": (mp, [mp]),
        "mat_from_array": (mp, [_f32p, C.c_size_t, C.c_size_t]),
        "medmad_normalise_array": (None, [_f32p, C.c_size_t]),
        "medianf": (C.c_float, [_f32p, C.c_size_t]),
        "madf": (C.c_float, [_f32p, C.c_size_t, _f32p]),
        "trim_and_segment_raw": (_RawTable, [_RawTable, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float]),
        "trim_raw_by_mad": (_RawTable, [_RawTable, C.c_size_t, C.c_float]),
        "get_raw_model": (C.c_int, [C.c_char_p]),
        "raw_model_string": (C.c_char_p, [C.c_int]),
        "get_raw_model_stride": (C.c_int, [C.c_int]),
        "get_raw_model_stride_from_string": (C.c_int, [C.c_char_p]),
        "get_posterior_function": (C.c_void_p, [C.c_int]),
        "decode_transducer": (C.c_float, [mp, C.c_float, C.c_float, C.c_float, _i32p, C.c_bool]),
        "decode_crf": (C.c_float, [mp, _i32p]),
        "posterior_crf": (mp, [mp]),
        "map_to_sequence_viterbi": (C.c_float, [mp, C.c_float, C.c_float, C.c_float, _i32p, C.c_size_t, _i32p]),
        "map_to_sequence_forward": (C.c_float, [mp, C.c_float, C.c_float, C.c_float, _i32p, C.c_size_t]),
        "map_to_sequence_viterbi_banded": (C.c_float, [mp, C.c_float, C.c_float, C.c_float, _i32p, C.c_size_t,
                                                       C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "map_to_sequence_forward_banded": (C.c_float, [mp, C.c_float, C.c_float, C.c_float, _i32p, C.c_size_t,
                                                       C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "are_bounds_sane": (C.c_bool, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t]),
        "encode_bases_to_integers": (C.c_void_p, [C.c_char_p, C.c_size_t, C.c_size_t]),
        "nanonet_posterior": (mp, [_EventTable, C.c_float, C.c_float, C.c_float, C.c_bool]),
        "nanonet_features_from_events": (mp, [_EventTable, C.c_bool]),
        "detect_events": (_EventTable, [_RawTable, DetectorParam]),
        "sb2_events_posterior_batch": (C.c_int, [C.c_void_p, C.POINTER(_EventTable), C.c_size_t, C.c_float, C.c_float,
                                                 C.c_float, C.c_bool, C.POINTER(mp)]),
        "sb2_batch_posterior_crf": (C.c_int, [C.c_void_p]),
        "sb2_batch_download_base_probs": (C.c_int, [C.c_void_p, C.c_size_t, _f32p]),
        "overlapper": (C.c_void_p, [_i32p, C.c_size_t, C.c_int, _i32p]),
        "crfpath_to_basecall": (C.c_void_p, [_i32p, C.c_size_t, _i32p]),
        "homopolymer_path": (C.c_int, [mp, _i32p, C.c_int]),
        "get_homopolymer_calculation": (C.c_int, [C.c_char_p]),
        "sb2_default_params": (Params, []),
        "sb2_engine_create": (C.c_void_p, [C.c_int, C.c_char_p]),
        "sb2_engine_destroy": (None, [C.c_void_p]),
        "sb2_engine_load_blob": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
        "sb2_last_error": (C.c_char_p, []),
        "sb2_engine_launch_count": (C.c_uint64, [C.c_void_p]),
        "sb2_engine_trim_pool": (C.c_int, [C.c_void_p]),
        "sb2_engine_realloc_count": (C.c_uint64, [C.c_void_p]),
        "sb2_batch_create": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.c_size_t]),
        "sb2_batch_destroy": (None, [C.c_void_p]),
        "sb2_batch_nblock": (C.c_size_t, [C.c_void_p, C.c_size_t]),
        "sb2_batch_total_blocks": (C.c_size_t, [C.c_void_p]),
        "sb2_batch_nstate": (C.c_size_t, [C.c_void_p]),
        "sb2_batch_upload": (C.c_int, [C.c_void_p, C.POINTER(_f32p)]),
        "sb2_batch_upload_concat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
        "sb2_batch_total_samples_padded": (C.c_size_t, [C.c_void_p]),
        "sb2_batch_sample_offset": (C.c_size_t, [C.c_void_p, C.c_size_t]),
        "sb2_host_alloc_pinned": (C.c_void_p, [C.c_size_t]),
        "sb2_host_free_pinned": (None, [C.c_void_p]),
        "sb2_batch_keep_layers": (C.c_int, [C.c_void_p, C.c_int]),
        "sb2_batch_forward": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_bool]),
        "sb2_batch_decode": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
        "sb2_batch_run": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
        "sb2_batch_sync": (C.c_int, [C.c_void_p]),
        "sb2_batch_download_posterior": (C.c_int, [C.c_void_p, C.c_size_t, _f32p, C.c_size_t]),
        "sb2_batch_download_paths": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
        "sb2_batch_download_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, _f32p]),
        "sb2_batch_time": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_int, C.c_int, _f32p, _f32p, _f32p]),
        "sb2_batch_stage_ms": (C.c_int, [C.c_void_p, _f32p, C.c_int]),
        "sb2_batch_stage_offsets": (C.c_int, [C.c_void_p, _f32p, C.c_int]),
        "sb2_basecall_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_f32p), C.POINTER(C.c_size_t), C.c_size_t,
                                         C.POINTER(Params), C.POINTER(_Call)]),
        "sb2_batch_basecall": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Params), C.POINTER(_Call)]),
        "sb2_calls_free": (None, [C.POINTER(_Call), C.c_size_t]),
        "sb2_default_trim": (Trim, []),
        "sb2_prepare_reads": (C.c_int, [C.c_void_p, C.POINTER(_f32p), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(Trim),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(_f32p)]),
        "sb2_basecall_raw_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_f32p), C.POINTER(C.c_size_t), C.c_size_t,
                                             C.POINTER(Trim), C.POINTER(Params), C.POINTER(_Call), C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t)]),
        "sb2_multi_time": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Params), C.c_int, C.c_int, _f32p]),
        "sb2_multi_stream_time": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Params), _i32p, _f32p]),
        "sb2_conv_plan_debug": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, _i32p, C.c_int]),
    }
    for name in MODELS:
        sig[_posterior_symbol(name)] = (mp, [_RawTable, C.c_float, C.c_float, C.c_float, C.c_bool])
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None     # filled by tests from include/scrappie_b200.h


def last_error():
    return lib().sb2_last_error().decode()


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _take_string(ptr):
    if not ptr:
        return None
    s = C.string_at(ptr).decode()
    _libc.free(ptr)
    return s


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def default_params(**kw):
    p = lib().sb2_default_params()
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown parameter %r" % k)
        setattr(p, k, v)
    return p


# ---------------------------------------------------------------------------
# scrappy-compatible single-read API
# ---------------------------------------------------------------------------

class RawTable(object):
    """A raw signal with trimming bounds (python/scrappy/__init__.py:47-111)."""

    def __init__(self, data, start=0, end=None):
        self._data = np.ascontiguousarray(data, dtype=np.float32).copy()
        if end is None:
            end = len(self._data)
        self._rt = _RawTable(None, len(self._data), start, end, _fp(self._data))

    def data(self, as_numpy=False):
        return self._data[self.start:self.end] if as_numpy else self._rt

    @property
    def start(self):
        return self._rt.start

    @property
    def end(self):
        return self._rt.end

    def trim(self, start=200, end=10, varseg_chunk=100, varseg_thresh=0.0):
        """Trim by MAD segmentation then fixed amounts (trim_and_segment_raw).  The C
        routine frees its buffer when nothing is left, so it works on a malloc'd copy."""
        if varseg_chunk < 2:
            # trim_raw_by_mad fails here WITHOUT freeing the buffer, while "nothing left" frees it: rejected before
            # the call so that the two cases cannot be confused (and nothing leaks)
            raise ValueError("varseg_chunk must be at least 2")
        _libc.malloc.restype = C.c_void_p
        _libc.malloc.argtypes = [C.c_size_t]
        buf = _libc.malloc(max(self._data.nbytes, 4))
        if not buf:
            raise MemoryError("trim: cannot allocate the working copy")
        C.memmove(buf, self._data.ctypes.data, self._data.nbytes)
        rt = _RawTable(None, self._rt.n, self._rt.start, self._rt.end, C.cast(buf, _f32p))
        out = lib().trim_and_segment_raw(rt, start, end, varseg_chunk, varseg_thresh)
        if not out.raw:
            raise RuntimeError("no signal left after trimming")
        _libc.free(buf)
        self._rt.start, self._rt.end = out.start, out.end
        return self

    def scale(self):
        """med-MAD normalise the trimmed part in place (medmad_normalise_array)."""
        view = self._data[self.start:self.end]
        lib().medmad_normalise_array(_fp(view), view.size)
        return self


class ScrappyMatrix(object):
    """Owner of a C scrappie_matrix (python/scrappy/__init__.py:147-192)."""

    def __init__(self, ptr):
        self._ptr = ptr

    def __del__(self):
        if getattr(self, "_ptr", None):
            lib().free_scrappie_matrix(self._ptr)
            self._ptr = None

    @property
    def shape(self):
        m = self._ptr.contents
        return int(m.nc), int(m.nr)

    def data(self, as_numpy=False, sloika=False):
        if not as_numpy:
            return self._ptr
        m = self._ptr.contents
        a = np.ctypeslib.as_array(m.f, shape=(m.nc, m.stride))[:, :m.nr]
        if sloika:
            a = np.hstack((a[:, m.nr - 1:m.nr], a[:, :m.nr - 1]))
        return np.array(a, dtype=np.float32, order="C")      # always a copy: the C matrix may be freed before its user is done

    def padded(self):
        m = self._ptr.contents
        return np.ctypeslib.as_array(m.f, shape=(m.nc, m.stride)).copy()

    @classmethod
    def from_numpy(cls, a, nr=None):
        """[nblock, >=nr] float32 -> scrappie_matrix with nr rows per column."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        nr = a.shape[1] if nr is None else nr
        packed = np.ascontiguousarray(a[:, :nr])
        return cls(lib().mat_from_array(_fp(packed), nr, a.shape[0]))


def _posterior_fn(model):
    if model not in MODELS:
        raise KeyError("Model type '{}' not recognised.".format(model))
    return getattr(lib(), _posterior_symbol(model))


def calc_post(rt, model='rgrgr_r94', min_prob=1e-6, log=True, tempW=1.0, tempb=1.0):
    if not log and model == 'rnnrf_r94':
        raise ValueError("Returning non-log transformed matrix not supported for model type 'rnnrf_r94'.")
    if not isinstance(rt, RawTable):
        raise TypeError('`rt` should be a RawTable.')
    ptr = _posterior_fn(model)(rt.data(), min_prob, tempW, tempb, log)
    if not ptr:
        raise RuntimeError('An unknown error occurred during posterior calculation. (%s)' % last_error())
    return ScrappyMatrix(ptr)


def _decode_post(post, stay_pen=0.0, skip_pen=0.0, local_pen=2.0, use_slip=False):
    nblock, nstate = post.shape
    path = np.zeros(nblock + 1, dtype=np.int32)
    score = lib().decode_transducer(post.data(), stay_pen, skip_pen, local_pen, _ip(path), use_slip)
    if score != score:
        raise RuntimeError("decode_transducer failed: %s" % last_error())
    pos = np.zeros(nblock + 1, dtype=np.int32)
    call = _take_string(lib().overlapper(_ip(path), nblock + 1, nstate - 1, _ip(pos)))
    return call, score, pos


def _decode_post_crf(post):
    nblock, nstate = post.shape
    path = np.zeros(nblock + 1, dtype=np.int32)
    score = lib().decode_crf(post.data(), _ip(path))
    if score != score:
        raise RuntimeError("decode_crf failed: %s" % last_error())
    pos = np.zeros(nblock + 1, dtype=np.int32)
    call = _take_string(lib().crfpath_to_basecall(_ip(path), nblock, _ip(pos)))
    return call, score, pos


def decode_post(post, model='rgrgr_r94', **kwargs):
    if not isinstance(post, ScrappyMatrix):
        raise TypeError('`post` should be a ScrappyMatrix.')
    if model not in MODELS:
        raise KeyError("Model type '{}' not recognised.".format(model))
    return _decode_post_crf(post, **kwargs) if model == 'rnnrf_r94' else _decode_post(post, **kwargs)


def decode_path(post, model='rgrgr_r94', stay_pen=0.0, skip_pen=0.0, local_pen=2.0, use_slip=False):
    """Viterbi path (nblock + 1 ints) and score straight from the C decoders."""
    nblock, _ = post.shape
    path = np.zeros(nblock + 1, dtype=np.int32)
    if model == 'rnnrf_r94':
        score = lib().decode_crf(post.data(), _ip(path))
    else:
        score = lib().decode_transducer(post.data(), stay_pen, skip_pen, local_pen, _ip(path), use_slip)
    return float(score), path


def get_model_stride(model):
    stride = lib().get_raw_model_stride_from_string(model.encode())
    if stride == -1:
        raise ValueError("Invalid scrappie model '{}'.".format(model))
    return stride


def detect_events(rt, param=None):
    """Segment a raw signal (a RawTable, trimmed range) into events: [n, 4] array of (start, length, mean, stdv)
    (detect_events, src/event_detection.c:270-320; host code)."""
    rt = rt if isinstance(rt, RawTable) else RawTable(rt)
    et = lib().detect_events(rt.data(), param or DetectorParam())
    if not et.event:
        raise RuntimeError("detect_events failed")
    out = np.array([[et.event[i].start, et.event[i].length, et.event[i].mean, et.event[i].stdv] for i in range(et.n)],
                   dtype=np.float64)
    _libc.free(C.cast(et.event, C.c_void_p))
    return out


def event_features(events):
    """Studentised event features [n, 4] (nanonet_features_from_events, src/nnfeatures.c:76-115; host)."""
    et = events if isinstance(events, EventTable) else EventTable(events)
    ptr = lib().nanonet_features_from_events(et.table, True)
    if not ptr:
        raise RuntimeError("nanonet_features_from_events failed")
    return ScrappyMatrix(ptr).data(as_numpy=True)


def calc_post_events(events, min_prob=1e-6, log=True, tempW=1.0, tempb=1.0):
    """nanonet_posterior (interface/scrappie.h:47-48): posterior of the events (LSTM) model as a ScrappyMatrix."""
    et = events if isinstance(events, EventTable) else EventTable(events)
    ptr = lib().nanonet_posterior(et.table, min_prob, tempW, tempb, log)
    if not ptr:
        raise RuntimeError("nanonet_posterior failed: %s" % last_error())
    return ScrappyMatrix(ptr)


def guess_state_properties(nstate):
    """(alphabet length, k-mer length) of a posterior with `nstate` rows (python/scrappy/__init__.py:25-44)."""
    for alpha_len in (4,):
        kmer_len = int(round(np.log(nstate - 1) / np.log(alpha_len)))
        if alpha_len ** kmer_len + 1 == nstate:
            return alpha_len, kmer_len
    raise ValueError("Cannot guess state properties from %d states." % nstate)


def encode_bases(sequence, kmer_len):
    """k-mer states of a base sequence (encode_bases_to_integers)."""
    b = sequence.encode()
    ptr = lib().encode_bases_to_integers(b, len(b), kmer_len)
    if not ptr:
        raise RuntimeError('An unknown error occurred whilst encoding sequence.')
    n = len(b) - kmer_len + 1
    out = np.ctypeslib.as_array(C.cast(ptr, _i32p), shape=(n,)).copy()
    _libc.free(C.c_void_p(ptr))
    return out


def map_post_to_sequence(post, sequence, stay_pen=0, skip_pen=0, local_pen=4.0, viterbi=False, path=False, bands=None):
    """Local-global alignment of a posterior to a base sequence, forward or Viterbi, optionally banded
    (python/scrappy/__init__.py:492-578; same arguments and return value: (score, path or None))."""
    if path and not viterbi:
        raise ValueError('Cannot calulate path with `viterbi==False`.')
    if not isinstance(post, ScrappyMatrix):
        raise TypeError('`post` should be a ScrappyMatrix.')
    nblock, nstate = post.shape
    alpha_len, kmer_len = guess_state_properties(nstate)
    seq = encode_bases(sequence, kmer_len)
    seq_len = seq.size
    path_data = np.zeros(nblock, dtype=np.int32) if (viterbi and path) else None
    sp = C.POINTER(C.c_size_t)
    if bands is None:
        if viterbi:
            score = lib().map_to_sequence_viterbi(post.data(), stay_pen, skip_pen, local_pen, _ip(seq), seq_len,
                                                  _ip(path_data) if path_data is not None else None)
        else:
            score = lib().map_to_sequence_forward(post.data(), stay_pen, skip_pen, local_pen, _ip(seq), seq_len)
    else:
        if isinstance(bands, int):
            gradient = seq_len / nblock
            hband = 2 * bands * gradient / 2
            bands = [np.ascontiguousarray(np.array(x, dtype=np.uintp)) for x in (
                [max(0, x * gradient - hband) for x in range(nblock)],
                [min(seq_len, x * gradient + hband) for x in range(nblock)])]
        elif len(bands) == 2:
            bands = [np.ascontiguousarray(x, dtype=np.uintp) for x in bands]
        else:
            raise ValueError('`bands` should be `None`, an integer, or length 2.')
        lo, hi = (x.ctypes.data_as(sp) for x in bands)
        if not lib().are_bounds_sane(lo, hi, nblock, seq_len):
            raise ValueError('Supplied banding structure is not valid.')
        func = lib().map_to_sequence_viterbi_banded if viterbi else lib().map_to_sequence_forward_banded
        score = func(post.data(), stay_pen, skip_pen, local_pen, _ip(seq), seq_len, lo, hi)
    if score != score:
        raise RuntimeError('An unknown error occurred during alignment.')
    return score, path_data


def posterior_crf(post):
    """Per-block base probabilities (ACGT-) of a CRF transition matrix as an (nblock + 1, 5) array
    (`lib.posterior_crf`, python/scrappy/__init__.py:424-428)."""
    ptr = lib().posterior_crf(post.data())
    if not ptr:
        raise RuntimeError("posterior_crf failed: %s" % last_error())
    return ScrappyMatrix(ptr).data(as_numpy=True)


def basecall_raw(data, model='rgrgr_r94', with_base_probs=False, **kwargs):
    """Trim, normalise, run the network and decode one read
    (python/scrappy/__init__.py:403-430).  Returns (call, score, pos, start, end, base_probs); the last item
    is None unless `with_base_probs` (rnnrf_r94 only)."""
    if with_base_probs and model != 'rnnrf_r94':
        raise ValueError("Base probabilities can only be returned for model 'rnnrf_r94'.")
    raw = RawTable(data)
    raw.trim().scale()
    post = calc_post(raw, model, log=True)
    seq, score, pos = decode_post(post, model, **kwargs)
    base_probs = posterior_crf(post) if with_base_probs else None
    return seq, score, pos, raw.start, raw.end, base_probs


# ---------------------------------------------------------------------------
# batch API
# ---------------------------------------------------------------------------

class Engine(object):
    """One CUDA device with resident model weights (sb2_engine)."""

    def __init__(self, device=0, weights_dir=None):
        wd = (weights_dir or WEIGHTS_DIR).encode()
        self._h = lib().sb2_engine_create(device, wd)
        if not self._h:
            raise RuntimeError("sb2_engine_create failed: %s" % last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().sb2_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def load_blob(self, model, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        rc = lib().sb2_engine_load_blob(self._h, _MODEL_ENUM[model], blob.ctypes.data, blob.size)
        if rc:
            raise RuntimeError("load_blob failed: %s" % last_error())

    @property
    def launches(self):
        return int(lib().sb2_engine_launch_count(self._h))

    @property
    def reallocs(self):
        return int(lib().sb2_engine_realloc_count(self._h))

    def trim_pool(self):
        """Free the idle workspaces basecall_batch keeps between calls; returns how many."""
        return int(lib().sb2_engine_trim_pool(self._h))

    def batch(self, model, nsample):
        return Batch(self, model, nsample)

    @staticmethod
    def _trim(trim=None, **kw):
        t = trim or lib().sb2_default_trim()
        for k, v in kw.items():
            setattr(t, k, v)
        return t

    def prepare_reads(self, raws, trim=None, normalise=True, **kw):
        """trim_and_segment_raw + medmad_normalise_array on the device for a list of untrimmed signals.
        Returns (start, end, normalised) -- `normalised[r]` is the scaled raws[r][start[r]:end[r]] (None when
        the read trims to nothing)."""
        t = self._trim(trim, **kw)
        sigs = [np.ascontiguousarray(s, dtype=np.float32) for s in raws]
        n = len(sigs)
        ptrs = (_f32p * n)(*[_fp(s) for s in sigs])
        lens = (C.c_size_t * n)(*[s.size for s in sigs])
        start = (C.c_size_t * n)()
        end = (C.c_size_t * n)()
        outs = [np.zeros(s.size, dtype=np.float32) for s in sigs] if normalise else None
        optrs = (_f32p * n)(*[_fp(o) for o in outs]) if normalise else None
        rc = lib().sb2_prepare_reads(self._h, ptrs, lens, n, C.byref(t), start, end, optrs)
        if rc:
            raise RuntimeError("sb2_prepare_reads failed: %s" % last_error())
        start, end = [int(x) for x in start], [int(x) for x in end]
        norm = None
        if normalise:
            norm = [o[:e - s_] if e > s_ else None for o, s_, e in zip(outs, start, end)]
        return start, end, norm

    def basecall_raw_batch(self, model, raws, params=None, trim=None, **kw):
        """calculate_post (src/scrappie_raw.c:265-315) for a list of untrimmed pA signals, everything on the
        device.  Returns a list of (bases or None, score, nblock, start, end)."""
        params = params or default_params()
        t = self._trim(trim, **kw)
        sigs = [np.ascontiguousarray(s, dtype=np.float32) for s in raws]
        n = len(sigs)
        ptrs = (_f32p * n)(*[_fp(s) for s in sigs])
        lens = (C.c_size_t * n)(*[s.size for s in sigs])
        start = (C.c_size_t * n)()
        end = (C.c_size_t * n)()
        out = (_Call * n)()
        rc = lib().sb2_basecall_raw_batch(self._h, _MODEL_ENUM[model], ptrs, lens, n, C.byref(t), C.byref(params), out,
                                          start, end)
        if rc < 0:
            raise RuntimeError("sb2_basecall_raw_batch failed: %s" % last_error())
        return [(_take_string(o.bases), float(o.score), int(o.nblock), int(s_), int(e))
                for o, s_, e in zip(out, start, end)]

    def events_posterior_batch(self, tables, min_prob=1e-6, log=True, tempW=1.0, tempb=1.0):
        """nanonet_posterior for a list of event tables ([n, 3] arrays or EventTable) in one pass; list of
        ScrappyMatrix (None where the reference would return NULL)."""
        ets = [t if isinstance(t, EventTable) else EventTable(t) for t in tables]
        n = len(ets)
        arr = (_EventTable * n)(*[e.table for e in ets])
        mp = C.POINTER(_Mat)
        out = (mp * n)()
        rc = lib().sb2_events_posterior_batch(self._h, arr, n, min_prob, tempW, tempb, log, out)
        if rc < 0:
            raise RuntimeError("sb2_events_posterior_batch failed: %s" % last_error())
        return [ScrappyMatrix(o) if o else None for o in out]

    def prepare_call(self, signals):
        """The argument arrays of sb2_basecall_batch for a list of (ordinary, pageable) float32 signals, built once so
        that a caller who basecalls the same buffers repeatedly pays no Python work per call."""
        sigs = [np.ascontiguousarray(s, dtype=np.float32) for s in signals]
        n = len(sigs)
        return (sigs, (_f32p * n)(*[_fp(s) for s in sigs]), (C.c_size_t * n)(*[s.size for s in sigs]), n)

    def basecall_prepared(self, model, prepared, params=None):
        """sb2_basecall_batch (the documented drop-in call: workspace from the engine's pool, signals staged from
        pageable memory) on arguments made by prepare_call.  Returns a CallSet."""
        params = params or default_params()
        _, ptrs, lens, n = prepared
        out = (_Call * n)()
        rc = lib().sb2_basecall_batch(self._h, _MODEL_ENUM[model], ptrs, lens, n, C.byref(params), out)
        if rc < 0:
            raise RuntimeError("sb2_basecall_batch failed: %s" % last_error())
        return CallSet(out, n)

    def basecall_batch(self, model, signals, params=None):
        """signals: list of trimmed + normalised float32 arrays.  Returns list of
        (bases, score, nblock)."""
        params = params or default_params()
        sigs = [np.ascontiguousarray(s, dtype=np.float32) for s in signals]
        n = len(sigs)
        ptrs = (_f32p * n)(*[_fp(s) for s in sigs])
        lens = (C.c_size_t * n)(*[s.size for s in sigs])
        out = (_Call * n)()
        rc = lib().sb2_basecall_batch(self._h, _MODEL_ENUM[model], ptrs, lens, n, C.byref(params), out)
        if rc < 0:
            raise RuntimeError("sb2_basecall_batch failed: %s" % last_error())
        return [(_take_string(o.bases), float(o.score), int(o.nblock)) for o in out]


class Batch(object):
    """Device workspace for a batch of reads (sb2_batch)."""

    def __init__(self, engine, model, nsample):
        self.engine = engine
        self.model = model
        self.nread = len(nsample)
        lens = (C.c_size_t * self.nread)(*[int(x) for x in nsample])
        self._h = lib().sb2_batch_create(engine._h, _MODEL_ENUM[model], lens, self.nread)
        if not self._h:
            raise RuntimeError("sb2_batch_create failed: %s" % last_error())
        L = lib()
        self.nblock = [int(L.sb2_batch_nblock(self._h, r)) for r in range(self.nread)]
        self.total_blocks = int(L.sb2_batch_total_blocks(self._h))
        self.nstate = int(L.sb2_batch_nstate(self._h))
        self.ostride = 4 * ((self.nstate + 3) // 4)
        self.sample_offset = [int(L.sb2_batch_sample_offset(self._h, r)) for r in range(self.nread)]
        self.total_samples_padded = int(L.sb2_batch_total_samples_padded(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().sb2_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc:
            raise RuntimeError("%s failed: %s" % (what, last_error()))

    def upload(self, signals):
        sigs = [np.ascontiguousarray(s, dtype=np.float32) for s in signals]
        ptrs = (_f32p * self.nread)(*[_fp(s) for s in sigs])
        self._check(lib().sb2_batch_upload(self._h, ptrs), "upload")

    def upload_concat(self, ptr, pinned_async=True):
        self._check(lib().sb2_batch_upload_concat(self._h, ptr, int(pinned_async)), "upload_concat")

    def keep_layers(self, keep=True):
        self._check(lib().sb2_batch_keep_layers(self._h, int(keep)), "keep_layers")

    def forward(self, params=None, return_log=True):
        params = params or default_params()
        self._check(lib().sb2_batch_forward(self._h, C.byref(params), return_log), "forward")

    def decode(self, params=None):
        params = params or default_params()
        self._check(lib().sb2_batch_decode(self._h, C.byref(params)), "decode")

    def run(self, params=None):
        """forward (log posterior) + decode; a captured CUDA graph is replayed from the third call on."""
        params = params or default_params()
        self._check(lib().sb2_batch_run(self._h, C.byref(params)), "run")

    def sync(self):
        self._check(lib().sb2_batch_sync(self._h), "sync")

    def posterior_crf(self):
        self._check(lib().sb2_batch_posterior_crf(self._h), "posterior_crf")

    def base_probs(self, read):
        out = np.zeros((self.nblock[read] + 1, 8), dtype=np.float32)
        self._check(lib().sb2_batch_download_base_probs(self._h, read, _fp(out)), "download_base_probs")
        return out[:, :5]

    def posterior(self, read):
        out = np.zeros((self.nblock[read], self.ostride), dtype=np.float32)
        self._check(lib().sb2_batch_download_posterior(self._h, read, _fp(out), self.ostride), "download_posterior")
        return out

    def layer(self, layer, read, H):
        out = np.zeros((self.nblock[read], H), dtype=np.float32)
        self._check(lib().sb2_batch_download_layer(self._h, layer, read, _fp(out)), "download_layer")
        return out

    def paths(self, out_paths=None, out_scores=None):
        """Returns (list of per-read path arrays, scores)."""
        paths = np.zeros(self.total_blocks + self.nread, dtype=np.int32) if out_paths is None else out_paths
        scores = np.zeros(self.nread, dtype=np.float32) if out_scores is None else out_scores
        self._check(lib().sb2_batch_download_paths(self._h, paths.ctypes.data, scores.ctypes.data), "download_paths")
        per_read, off = [], 0
        for r in range(self.nread):
            per_read.append(paths[off:off + self.nblock[r] + 1])
            off += self.nblock[r] + 1
        return per_read, scores

    def time(self, params=None, nrep=1, flush_l2=True):
        """CUDA-event timing of forward+decode on the batch's stream: (total, forward, decode) ms arrays."""
        params = params or default_params()
        tot = np.zeros(nrep, dtype=np.float32)
        fwd = np.zeros(nrep, dtype=np.float32)
        dec = np.zeros(nrep, dtype=np.float32)
        self._check(lib().sb2_batch_time(self._h, C.byref(params), nrep, int(flush_l2), _fp(tot), _fp(fwd), _fp(dec)), "time")
        return tot, fwd, dec

    def basecall(self, concat_ptr=None, pinned=True, params=None, lazy=False):
        """Upload (optional), forward, decode, download, homopolymer, overlapper on this workspace.
        Returns list of (bases, score, nblock); with lazy=True a CallSet that leaves the base strings
        in C memory until they are asked for (no per-read Python work inside the call)."""
        params = params or default_params()
        out = (_Call * self.nread)()
        rc = lib().sb2_batch_basecall(self._h, concat_ptr, int(pinned), C.byref(params), out)
        if rc < 0:
            raise RuntimeError("sb2_batch_basecall failed: %s" % last_error())
        cs = CallSet(out, self.nread)
        return cs if lazy else cs.tolist()

    STAGES = ("conv", "affine1", "scan1", "affine2", "scan2", "affine3", "scan3", "affine4", "scan4",
              "affine5", "scan5", "head_gemm", "head_finish", "decode")

    def stage_ms(self):
        buf = np.zeros(len(self.STAGES), dtype=np.float32)
        lib().sb2_batch_stage_ms(self._h, _fp(buf), buf.size)
        return dict(zip(self.STAGES, [float(x) for x in buf]))

    def stage_offsets(self):
        """Stage boundaries of the last `multi_time` run: ms after the first batch's first launch."""
        buf = np.zeros(len(self.STAGES) + 1, dtype=np.float32)
        lib().sb2_batch_stage_offsets(self._h, _fp(buf), buf.size)
        return [float(x) for x in buf]


class CallSet(object):
    """The sb2_call records of one batch.  Scores / lengths are numpy views of the C array; base strings
    are converted on access and freed with the set."""
    _DT = np.dtype({"names": ["bases", "score", "nblock", "nbase"], "formats": ["u8", "f4", "u8", "u8"],
                    "offsets": [0, 8, 16, 24], "itemsize": 32})

    def __init__(self, calls, n):
        assert C.sizeof(_Call) == 32
        self._calls, self._n = calls, n
        self._view = np.frombuffer(calls, dtype=self._DT, count=n)

    def __len__(self):
        return self._n

    @property
    def scores(self):
        return self._view["score"]

    @property
    def nbase(self):
        return self._view["nbase"]

    @property
    def nblock(self):
        return self._view["nblock"]

    def bases(self, i):
        p = self._calls[i].bases
        return C.string_at(p).decode() if p else None

    def __getitem__(self, i):
        return self.bases(i), float(self._view["score"][i]), int(self._view["nblock"][i])

    def tolist(self):
        out = [self[i] for i in range(self._n)]
        self.close()
        return out

    def close(self):
        if getattr(self, "_calls", None) is not None:
            self._view = None
            lib().sb2_calls_free(self._calls, self._n)
            self._calls = None

    __del__ = close


def multi_time(batches, params=None, nrep=1, flush_l2=True):
    """CUDA-event time (ms per repetition) of forward+decode over several concurrent batches."""
    params = params or default_params()
    hs = (C.c_void_p * len(batches))(*[b._h for b in batches])
    ms = np.zeros(nrep, dtype=np.float32)
    rc = lib().sb2_multi_time(hs, len(batches), C.byref(params), nrep, int(flush_l2), _fp(ms))
    if rc:
        raise RuntimeError("sb2_multi_time failed: %s" % last_error())
    return ms


def multi_stream_time(batches, params=None, nrep=1):
    """Device time (ms, total) of `nrep` forward+decode steps per batch (an int, or one count per batch), the
    batches free-running on their own streams with no synchronisation between steps (streaming throughput)."""
    params = params or default_params()
    hs = (C.c_void_p * len(batches))(*[b._h for b in batches])
    ms = np.zeros(1, dtype=np.float32)
    reps = np.ascontiguousarray(np.broadcast_to(np.asarray(nrep, dtype=np.int32), (len(batches),)))
    rc = lib().sb2_multi_stream_time(hs, len(batches), C.byref(params), _ip(reps), _fp(ms))
    if rc:
        raise RuntimeError("sb2_multi_stream_time failed: %s" % last_error())
    return float(ms[0])


_caller = None


def caller_lib():
    """libsb2_caller.so: the C caller of the batch drop-in call (examples/batch_caller.c)."""
    global _caller
    if _caller is None:
        lib()                                           # libscrappie_b200.so first: the caller links against it
        L = C.CDLL(os.path.join(_HERE, "libsb2_caller.so"))
        L.sb2_caller_run.restype = C.c_double
        L.sb2_caller_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(_f32p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                     C.c_int, _i32p, C.c_int, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p),
                                     _f32p, C.POINTER(C.c_size_t)]
        L.sb2_caller_free.argtypes = [C.c_void_p]
        L.sb2_caller_last_error.restype = C.c_char_p
        _caller = L
    return _caller


class CallerJob(object):
    """A workload prepared for sb2_caller_run: batches of reads in ordinary host memory, handed to a team of C host
    threads that each call sb2_basecall_batch (the documented drop-in call) on the next batch -- what `scrappie raw`'s
    OpenMP loop would do with the GPU library behind it.  No Python runs while the job does."""

    def __init__(self, engine, model, groups, order=None):
        self.engine, self.model = engine, model
        self.sigs = [np.ascontiguousarray(s, dtype=np.float32) for g in groups for s in g]
        n = len(self.sigs)
        self.nread = n
        self.ptrs = (_f32p * n)(*[_fp(s) for s in self.sigs])
        self.lens = (C.c_size_t * n)(*[s.size for s in self.sigs])
        starts = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(np.uintp)
        self.starts = np.ascontiguousarray(starts)
        self.nbatch = len(groups)
        self.order = None if order is None else np.ascontiguousarray(order, dtype=np.int32)

    def run(self, nstep, nthread, params=None, want_bases=False):
        """Returns (seconds, bases called in the last step, [base strings of the last step])."""
        params = params or default_params()
        out = (C.c_void_p * self.nread)() if want_bases else None
        scores = np.zeros(self.nread, dtype=np.float32)
        nb = C.c_size_t(0)
        secs = caller_lib().sb2_caller_run(self.engine._h, _MODEL_ENUM[self.model], self.ptrs, self.lens,
                                           self.starts.ctypes.data_as(C.POINTER(C.c_size_t)), self.nbatch,
                                           _ip(self.order) if self.order is not None else None, int(nstep), int(nthread),
                                           C.byref(params), out, _fp(scores), C.byref(nb))
        if secs < 0:
            raise RuntimeError("sb2_caller_run failed: %s" % caller_lib().sb2_caller_last_error().decode())
        bases = None
        if want_bases:
            bases = []
            for p in out:
                bases.append(C.string_at(p).decode() if p else None)
                if p:
                    caller_lib().sb2_caller_free(p)
        return secs, int(nb.value), bases, scores


class PinnedBuffer(object):
    """Page-locked host float32 buffer (cudaMallocHost) exposed as a numpy array."""

    def __init__(self, nfloat):
        self.ptr = lib().sb2_host_alloc_pinned(nfloat * 4)
        if not self.ptr:
            raise RuntimeError("pinned allocation failed")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _f32p), shape=(nfloat,))

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().sb2_host_free_pinned(self.ptr)
            self.ptr = None

    __del__ = close


def conv_plan(nsample, winlen, stride):
    """Host-side convolution tail plan as {column: [(x0, tap0, ntap), ...]} plus (first_col, ncol)."""
    buf = np.zeros(2 + 24 * (1 + 9), dtype=np.int32)
    rc = lib().sb2_conv_plan_debug(nsample, winlen, stride, _ip(buf), buf.size)
    if rc:
        raise RuntimeError("conv plan failed: %s" % last_error())
    first, ncol = int(buf[0]), int(buf[1])
    plan = {}
    for c in range(24):
        base = 2 + c * 10
        nseg = int(buf[base])
        if first + c < ncol:
            plan[first + c] = [tuple(int(v) for v in buf[base + 1 + 3 * s: base + 4 + 3 * s]) for s in range(nseg)]
    return first, ncol, plan
