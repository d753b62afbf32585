#!/usr/bin/env python
"""Extract model weights from the reference build into flat blobs.

The reference keeps its weights as non-static `_Mat` globals (hex-float C arrays
in src/models/*.h, pulled into src/networks.c:2-15).  They are therefore exported
data symbols of oracle/_ref/libscrappie_ref.so; this script reads them through
ctypes and writes one little-endian blob per model under scrappie_b200/weights/.

The blobs are DATA derived from nanoporetech/scrappie (MPL-2.0, see
scrappie_b200/weights/NOTICE); no reference source code is copied.

Blob layout (all little-endian):
    char     magic[8]   = "SB2WTS01"
    uint32   n_tensor
    uint32   conv_stride
    uint32   conv_act      0 = ELU, 1 = tanh       (src/networks.c:260 / :358)
    uint32   head          0 = softmax, 1 = globalnorm (src/networks.c:287 / :609)
    uint32   residual      1 = GRU layers wrapped in residual (src/networks.c:583)
    uint32   arch          0 = conv + 5 alternating unidirectional GRU layers + head (rgrgr / rnnrf),
                           1 = raw_r94: conv + 2 x (bidirectional GRU pair + feedforward2_tanh) + head
                               (src/networks.c:196-247); tensors gru1..4 = F1, B1, F2, B2, comb1/2 = FF1/FF2
                           2 = nanonet events model (src/networks.c:146-194): no convolution; 2 x (bidirectional LSTM
                               pair + feedforward2_tanh) + head; tensors gru1..4 = lstm F1, B1, F2, B2 with
                               sW = [H][4H] recurrent weights and sW2 = the 3H peephole weights `p`
    uint32   reserved[2]
    n_tensor x { char name[24]; uint32 nr, nc, stride, offset }   offset in floats
    float    data[]      each tensor exactly as the reference stores it:
                         column-major, nc columns of `stride` floats
"""
import ctypes
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class Mat(ctypes.Structure):
    _fields_ = [("nr", ctypes.c_size_t), ("nrq", ctypes.c_size_t), ("nc", ctypes.c_size_t),
                ("stride", ctypes.c_size_t), ("data", ctypes.POINTER(ctypes.c_float))]


LAYERS = ["gruB1", "gruF2", "gruB3", "gruF4", "gruB5"]

MODELS = {
    # name: (conv_act, head, residual)
    "rgrgr_r94": (0, 0, 0),
    "rgrgr_r941": (0, 0, 0),
    "rgrgr_r10": (1, 0, 0),
    "rnnrf_r94": (0, 1, 1),
}


def read_mat(lib, sym):
    m = Mat.in_dll(lib, sym)
    n = m.stride * m.nc
    arr = np.ctypeslib.as_array(m.data, shape=(n,)).astype("<f4").copy()
    assert m.stride == 4 * m.nrq, (sym, m.stride, m.nrq)
    return int(m.nr), int(m.nc), int(m.stride), arr


def extract_raw_r94(lib, outdir):
    """nanonet_raw_posterior's weights (src/networks.c:196-247, src/models/raw_r94.h)."""
    tensors = [("conv_W",) + read_mat(lib, "_conv_raw_W"), ("conv_b",) + read_mat(lib, "_conv_raw_b")]
    for i, lay in enumerate(["gruF1", "gruB1", "gruF2", "gruB2"], 1):
        for part in ("iW", "b", "sW", "sW2"):
            tensors.append(("gru%d_%s" % (i, part),) + read_mat(lib, "_%s_raw_%s" % (lay, part)))
    for i in (1, 2):
        for part in ("Wf", "Wb", "b"):
            tensors.append(("comb%d_%s" % (i, part),) + read_mat(lib, "_FF%d_raw_%s" % (i, part)))
    tensors.append(("FF_W",) + read_mat(lib, "_FF3_raw_W"))
    tensors.append(("FF_b",) + read_mat(lib, "_FF3_raw_b"))
    stride = ctypes.c_int.in_dll(lib, "conv_raw_stride").value
    write_blob("raw_r94", tensors, stride, 1, 0, 0, 1, outdir)


def extract_events(lib, outdir):
    """nanonet_posterior's weights (src/networks.c:146-194, src/models/nanonet_events.h)."""
    tensors = []
    for i, lay in enumerate(["lstmF1", "lstmB1", "lstmF2", "lstmB2"], 1):
        for part, sym in (("iW", "iW"), ("b", "b"), ("sW", "sW"), ("sW2", "p")):
            tensors.append(("gru%d_%s" % (i, part),) + read_mat(lib, "_%s_%s" % (lay, sym)))
    for i in (1, 2):
        for part in ("Wf", "Wb", "b"):
            tensors.append(("comb%d_%s" % (i, part),) + read_mat(lib, "_FF%d_%s" % (i, part)))
    tensors.append(("FF_W",) + read_mat(lib, "_FF3_W"))
    tensors.append(("FF_b",) + read_mat(lib, "_FF3_b"))
    write_blob("nanonet_events", tensors, 1, 0, 0, 0, 2, outdir)


def write_blob(model, tensors, stride, conv_act, head, residual, arch, outdir):
    hdr = b"SB2WTS01" + struct.pack("<8I", len(tensors), stride, conv_act, head, residual, arch, 0, 0)
    table = b""
    off = 0
    for name, nr, nc, st, arr in tensors:
        table += struct.pack("<24s4I", name.encode(), nr, nc, st, off)
        off += arr.size
    path = os.path.join(outdir, model + ".bin")
    with open(path, "wb") as fh:
        fh.write(hdr)
        fh.write(table)
        for _, _, _, _, arr in tensors:
            fh.write(arr.tobytes())
    print("%s: %d tensors, %d floats, conv stride %d -> %s" % (model, len(tensors), off, stride, path))


def extract(lib, model, outdir):
    conv_act, head, residual = MODELS[model]
    tensors = []
    tensors.append(("conv_W",) + read_mat(lib, "_conv_%s_W" % model))
    tensors.append(("conv_b",) + read_mat(lib, "_conv_%s_b" % model))
    for i, lay in enumerate(LAYERS, 1):
        for part in ("iW", "b", "sW", "sW2"):
            tensors.append(("gru%d_%s" % (i, part),) + read_mat(lib, "_%s_%s_%s" % (lay, model, part)))
    tensors.append(("FF_W",) + read_mat(lib, "_FF_%s_W" % model))
    tensors.append(("FF_b",) + read_mat(lib, "_FF_%s_b" % model))
    stride = ctypes.c_int.in_dll(lib, "conv_%s_stride" % model).value

    write_blob(model, tensors, stride, conv_act, head, residual, 0, outdir)


def main():
    so = os.path.join(ROOT, "oracle", "_ref", "libscrappie_ref.so")
    if not os.path.exists(so):
        sys.exit("build oracle/_ref first (make -C oracle ref)")
    lib = ctypes.CDLL(so)
    outdir = os.path.join(ROOT, "scrappie_b200", "weights")
    os.makedirs(outdir, exist_ok=True)
    for model in MODELS:
        extract(lib, model, outdir)
    extract_raw_r94(lib, outdir)
    extract_events(lib, outdir)


if __name__ == "__main__":
    main()
