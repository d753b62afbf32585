"""Minimal read-only HDF5 subset reader for nanopore fast5 files (no libhdf5 / h5py).

Only what the bundled single-read fast5 files use: superblock v0, v1 object
headers (with continuation blocks), symbol-table groups (v1 B-tree + local heap),
v1 attribute messages, data layout v3 (contiguous or chunked with a v1 chunk
B-tree) and the deflate filter.  Mirrors what the reference obtains through
libhdf5 in src/fast5_interface.c:130-217 (`read_raw`): the first read under
/Raw/Reads, its int16 `Signal` dataset and `read_id` attribute, and the float
attributes digitisation / offset / range of /UniqueGlobalKey/channel_id.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert self.b[8] == 0, "only superblock v0 supported"
        assert self.b[13] == 8 and self.b[14] == 8
        # root group symbol table entry starts at byte 56: link name off, obj header addr
        self.root = struct.unpack_from("<Q", self.b, 56 + 8)[0]

    # ---- object headers -------------------------------------------------
    def messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # ---- groups ---------------------------------------------------------
    def children(self, addr):
        b = self.b
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", b, body)
                assert b[heap:heap + 4] == b"HEAP"
                heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
                res = {}
                self._walk_group_btree(btree, heap_data, res)
                return res
        return {}

    def _walk_group_btree(self, addr, heap_data, res):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 0
        pos = addr + 24
        for i in range(nent):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
            if level > 0:
                self._walk_group_btree(child, heap_data, res)
            else:
                assert b[child:child + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                for s in range(nsym):
                    e = child + 8 + s * 40
                    name_off, ohdr = struct.unpack_from("<QQ", b, e)
                    p = heap_data + name_off
                    name = b[p:b.index(b"\0", p)].decode()
                    res[name] = ohdr

    def lookup(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            addr = self.children(addr)[part]
        return addr

    # ---- attributes -------------------------------------------------------
    @staticmethod
    def _pad8(n):
        return (n + 7) & ~7

    def attrs(self, addr):
        b = self.b
        out = {}
        for mtype, body, _ in self.messages(addr):
            if mtype != 0x0C:
                continue
            ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", b, body)
            assert ver == 1
            p = body + 8
            name = b[p:p + nsz].split(b"\0")[0].decode()
            p += self._pad8(nsz)
            dt = b[p:p + tsz]
            p += self._pad8(tsz)
            p += self._pad8(ssz)
            cls = dt[0] & 0x0F
            size = struct.unpack_from("<I", dt, 4)[0]
            raw = b[p:p + size]
            if cls == 1:
                out[name] = struct.unpack("<d" if size == 8 else "<f", raw)[0]
            elif cls == 0:
                signed = bool(dt[1] & 0x08)
                out[name] = int.from_bytes(raw, "little", signed=signed)
            elif cls == 3:
                out[name] = raw.split(b"\0")[0].decode()
            else:
                out[name] = raw
        return out

    # ---- datasets ---------------------------------------------------------
    def dataset_i16(self, addr):
        b = self.b
        dims = None
        layout = None
        deflate = False
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x01:
                ver, rank = b[body], b[body + 1]
                assert ver == 1
                dims = struct.unpack_from("<%dQ" % rank, b, body + 8)
            elif mtype == 0x03:
                assert (b[body] & 0x0F) == 0 and struct.unpack_from("<I", b, body + 4)[0] == 2
            elif mtype == 0x0B:
                deflate = True
            elif mtype == 0x08:
                ver, cls = b[body], b[body + 1]
                assert ver == 3
                if cls == 1:
                    layout = ("contig",) + struct.unpack_from("<QQ", b, body + 2)
                elif cls == 2:
                    rank = b[body + 2]
                    btree = struct.unpack_from("<Q", b, body + 3)[0]
                    cdims = struct.unpack_from("<%dI" % rank, b, body + 11)
                    layout = ("chunked", btree, cdims)
                else:
                    raise ValueError("compact layout unsupported")
        n = dims[0]
        if layout[0] == "contig":
            return np.frombuffer(b, dtype="<i2", count=n, offset=layout[1]).copy()
        out = np.zeros(n, dtype="<i2")
        chunks = []
        self._walk_chunk_btree(layout[1], len(layout[2]), chunks)
        clen = layout[2][0]
        for size, mask, off0, caddr in chunks:
            raw = b[caddr:caddr + size]
            if deflate and not (mask & 1):
                raw = zlib.decompress(raw)
            data = np.frombuffer(raw, dtype="<i2")
            m = min(clen, n - off0)
            out[off0:off0 + m] = data[:m]
        return out

    def _walk_chunk_btree(self, addr, rank, chunks):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 1
        keysz = 8 + 8 * rank
        pos = addr + 24
        for i in range(nent):
            k = pos + i * (keysz + 8)
            size, mask = struct.unpack_from("<II", b, k)
            off0 = struct.unpack_from("<Q", b, k + 8)[0]
            child = struct.unpack_from("<Q", b, k + keysz)[0]
            if level > 0:
                self._walk_chunk_btree(child, rank, chunks)
            else:
                chunks.append((size, mask, off0, child))


def read_raw(path, scale=True):
    """Equivalent of the reference's read_raw (src/fast5_interface.c:130-217).

    Returns dict(raw=float32 array (pA if scale) , read_id, digitisation, offset, range).
    Scaling is done in float32 exactly as the reference does
    (src/fast5_interface.c:196-202): (float(raw) + offset) * (range / digitisation).
    """
    f = H5File(path)
    reads = f.children(f.lookup("/Raw/Reads"))
    first = sorted(reads)[0] if len(reads) > 1 else next(iter(reads))
    raddr = reads[first]
    sig = f.dataset_i16(f.children(raddr)["Signal"])
    rattrs = f.attrs(raddr)
    ch = f.attrs(f.lookup("/UniqueGlobalKey/channel_id"))
    dig = np.float32(ch["digitisation"])
    off = np.float32(ch["offset"])
    rng = np.float32(ch["range"])
    out = dict(signal_i16=sig, read_id=rattrs.get("read_id", ""), digitisation=float(dig),
               offset=float(off), range=float(rng))
    if scale:
        unit = np.float32(rng / dig)
        out["raw"] = ((sig.astype(np.float32) + off) * unit).astype(np.float32)
    return out


if __name__ == "__main__":
    import sys
    for p in sys.argv[1:]:
        r = read_raw(p)
        print(p, len(r["raw"]), r["read_id"], r["digitisation"], r["offset"], r["range"], r["raw"][:4])
