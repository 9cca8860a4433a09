"""TEST INFRASTRUCTURE: a small, strict reader of the HDF5 file format subset `kallisto` writes (superblock version 0,
symbol-table groups, version-1 object headers, 1-D chunked datasets with the deflate filter, fixed-point / IEEE float /
fixed-length string types), written from the HDF5 File Format Specification (version 1.1 structures) independently of
csrc/h5_writer.hpp.  There is no libhdf5 / h5py in this image; this reader is what the tests hold the writer to, and it
checks every field a conforming reader depends on (signatures, versions, sizes, alignment, sorted names, B-tree keys,
addresses inside the file)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


def _need(cond, msg):
    if not cond:
        raise H5Error(msg)


class File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        _need(b[:8] == b"\x89HDF\r\n\x1a\n", "signature")
        sb_ver, fs_ver, root_ver, r0, shm_ver, so, sl, r1 = b[8:16]
        _need((sb_ver, fs_ver, root_ver, r0, shm_ver, r1) == (0, 0, 0, 0, 0, 0), "superblock version fields")
        _need((so, sl) == (8, 8), "8-byte offsets and lengths")
        self.leaf_k, self.int_k, flags = struct.unpack_from("<HHI", b, 16)
        _need(self.leaf_k >= 1 and self.int_k >= 1 and flags == 0, "B-tree K values / consistency flags")
        base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
        _need(base == 0 and free == UNDEF and driver == UNDEF, "base / free-space / driver addresses")
        _need(eof == len(b), "end-of-file address %d != file size %d" % (eof, len(b)))
        name_off, hdr, cache, rsv, bt, heap = struct.unpack_from("<QQIIQQ", b, 56)
        _need(name_off == 0 and cache == 1 and rsv == 0, "root symbol table entry")
        self.root = self._group(hdr, (bt, heap))

    # ---- low level ----
    def _at(self, addr, n):
        _need(addr != UNDEF and addr % 8 == 0 and addr + n <= len(self.b), "address %d (+%d) outside the file or unaligned" % (addr, n))
        return self.b[addr:addr + n]

    def _messages(self, addr):
        """version-1 object header -> list of (type, flags, data)"""
        ver, rsv, nmsg, refc, size = struct.unpack_from("<BBHII", self._at(addr, 16))
        _need(ver == 1 and rsv == 0 and refc == 1, "object header prefix")
        _need(size % 8 == 0, "object header size not a multiple of 8")
        body = self._at(addr + 16, size)
        out, o = [], 0
        while o < size:
            t, sz, fl = struct.unpack_from("<HHB", body, o)
            _need(sz % 8 == 0 and o + 8 + sz <= size, "message size")
            _need(body[o + 5:o + 8] == b"\0\0\0", "message reserved bytes")
            out.append((t, fl, body[o + 8:o + 8 + sz]))
            o += 8 + sz
        _need(len(out) == nmsg, "number of header messages %d != %d" % (len(out), nmsg))
        return out

    def _heap(self, addr):
        h = self._at(addr, 32)
        _need(h[:4] == b"HEAP" and h[4] == 0 and h[5:8] == b"\0\0\0", "local heap header")
        size, free, data = struct.unpack_from("<QQQ", h, 8)
        _need(free == 1 or free + 16 <= size, "local heap free list")       # 1 = H5HL_FREE_NULL
        return self._at(data, size)

    @staticmethod
    def _name(heap, off):
        _need(off < len(heap), "name offset outside the heap")
        end = heap.index(b"\0", off)
        return heap[off:end].decode()

    def _group(self, hdr, cached=None):
        msgs = self._messages(hdr)
        st = [m for m in msgs if m[0] == 0x0011]
        _need(len(st) == 1, "group without a symbol table message")
        bt, heap_addr = struct.unpack_from("<QQ", st[0][2])
        if cached is not None:
            _need(cached == (bt, heap_addr), "cached B-tree / heap addresses differ from the symbol table message")
        heap = self._heap(heap_addr)
        _need(heap[:1] == b"\0", "the heap must start with the empty name")
        node_size = 24 + (2 * self.int_k + 1) * 8 + 2 * self.int_k * 8
        node = self._at(bt, node_size)
        _need(node[:4] == b"TREE" and node[4] == 0 and node[5] == 0, "group B-tree node (type 0, level 0)")
        used, left, right = struct.unpack_from("<HQQ", node, 6)
        _need(left == UNDEF and right == UNDEF and used <= 2 * self.int_k, "group B-tree siblings / entries")
        keys = [struct.unpack_from("<Q", node, 24 + 16 * i)[0] for i in range(used + 1)]
        kids = [struct.unpack_from("<Q", node, 32 + 16 * i)[0] for i in range(used)]
        _need(self._name(heap, keys[0]) == "", "first key must be the empty name")
        entries = {}
        order = []
        for c, child in enumerate(kids):
            sn = self._at(child, 8 + 2 * self.leaf_k * 40)
            _need(sn[:4] == b"SNOD" and sn[4] == 1 and sn[5] == 0, "symbol table node")
            (n,) = struct.unpack_from("<H", sn, 6)
            _need(1 <= n <= 2 * self.leaf_k, "symbols in a node")
            last = None
            for i in range(n):
                no, oh, cache, rsv, s0, s1 = struct.unpack_from("<QQIIQQ", sn, 8 + 40 * i)
                _need(rsv == 0 and cache in (0, 1), "symbol table entry")
                nm = self._name(heap, no)
                order.append(nm)
                entries[nm] = (oh, (s0, s1) if cache == 1 else None)
                last = nm
            lo, hi = self._name(heap, keys[c]), self._name(heap, keys[c + 1])
            _need(hi == last, "B-tree key %d is not the largest name of its child" % (c + 1))
            _need(all(lo.encode() < x.encode() <= hi.encode() for x in order[-n:]), "names outside their key interval")
        _need([x.encode() for x in order] == sorted(x.encode() for x in order) and len(set(order)) == len(order), "names not sorted / unique")
        out = {}
        for nm, (oh, cached_sub) in entries.items():
            if cached_sub is not None or any(m[0] == 0x0011 for m in self._messages(oh)):
                out[nm] = self._group(oh, cached_sub)
            else:
                out[nm] = self._dataset(oh)
        return out

    def _dataset(self, hdr):
        msgs = {}
        for t, fl, d in self._messages(hdr):
            _need(t not in msgs, "duplicate message")
            msgs[t] = d
        for t in (0x0001, 0x0003, 0x0008):
            _need(t in msgs, "dataset without message %#x" % t)
        d = msgs[0x0001]                                   # dataspace, version 1
        _need(d[0] == 1 and d[1] == 1 and d[2] in (0, 1) and d[3:8] == b"\0" * 5, "dataspace message")
        (n,) = struct.unpack_from("<Q", d, 8)
        if d[2] == 1:
            _need(struct.unpack_from("<Q", d, 16)[0] == n, "maximum dimension")
        d = msgs[0x0003]                                   # datatype, version 1
        cls, ver = d[0] & 15, d[0] >> 4
        (size,) = struct.unpack_from("<I", d, 4)
        _need(ver == 1, "datatype version")
        if cls == 0:
            _need(d[1] == 0x08 and d[2] == 0 and d[3] == 0 and size == 4 and struct.unpack_from("<HH", d, 8) == (0, 32), "32-bit signed LE integer")
            dt = np.dtype("<i4")
        elif cls == 1:
            _need(d[1] == 0x20 and d[2] == 63 and d[3] == 0 and size == 8, "IEEE double flags")
            _need(struct.unpack_from("<HHBBBBI", d, 8) == (0, 64, 52, 11, 0, 52, 1023), "IEEE double properties")
            dt = np.dtype("<f8")
        elif cls == 3:
            _need(d[1] == 0 and d[2] == 0 and d[3] == 0 and size >= 1, "NUL-terminated ASCII string")
            dt = np.dtype("S%d" % size)
        else:
            raise H5Error("datatype class %d" % cls)
        if 0x0005 in msgs:                                 # fill value, version 2
            d = msgs[0x0005]
            _need(d[0] == 2 and d[1] in (1, 2, 3) and d[2] in (0, 1, 2) and d[3] in (0, 1), "fill value message")
            if d[3] == 1:
                _need(struct.unpack_from("<I", d, 4)[0] == 0, "fill value size")
        d = msgs[0x0008]                                   # data layout, version 3
        _need(d[0] == 3, "layout version")
        if d[1] == 1:
            addr, nbytes = struct.unpack_from("<QQ", d, 2)
            _need(nbytes == n * dt.itemsize, "contiguous size")
            raw = self._at(addr, nbytes)
        else:
            _need(d[1] == 2 and d[2] == 2, "chunked layout of rank 1")
            bt, cdim, esize = struct.unpack_from("<QII", d, 3)
            _need(cdim == n and esize == dt.itemsize, "one chunk of the whole vector")
            level = None
            if 0x000B in msgs:                             # filter pipeline, version 1
                p = msgs[0x000B]
                _need(p[0] == 1 and p[1] == 1 and p[2:8] == b"\0" * 6, "filter pipeline header")
                fid, nlen, fflags, ncd = struct.unpack_from("<HHHH", p, 8)
                _need(fid == 1 and nlen % 8 == 0 and ncd == 1 and fflags in (0, 1), "deflate filter")
                if nlen:
                    _need(p[16:16 + nlen].rstrip(b"\0") == b"deflate", "filter name")
                (level,) = struct.unpack_from("<I", p, 16 + nlen)
                _need(len(p) >= 16 + nlen + 8, "client data padding")
            key = 8 + 2 * 8
            node = self._at(bt, 24 + (2 * 32 + 1) * key + 2 * 32 * 8)
            _need(node[:4] == b"TREE" and node[4] == 1 and node[5] == 0, "chunk B-tree node (type 1, level 0)")
            used, left, right = struct.unpack_from("<HQQ", node, 6)
            _need(used == 1 and left == UNDEF and right == UNDEF, "one chunk")
            csize, mask, o0, o1 = struct.unpack_from("<IIQQ", node, 24)
            (caddr,) = struct.unpack_from("<Q", node, 24 + key)
            _, _, e0, e1 = struct.unpack_from("<IIQQ", node, 24 + key + 8)
            _need((mask, o0, o1) == (0, 0, 0) and e0 >= n and e1 == 0, "chunk keys")
            raw = self._at(caddr, csize)
            if level is not None:
                raw = zlib.decompress(raw)
                self.deflate_level = level
            _need(len(raw) == n * dt.itemsize, "chunk size after the filter")
        a = np.frombuffer(raw, dtype=dt)
        if cls == 3:
            _need(all(raw[(i + 1) * size - 1] == 0 for i in range(n)), "every string must keep its terminator")
            return [x.decode() for x in a.tolist()]          # numpy drops the padding NULs
        return a.copy()


def read(path):
    """-> nested dict {name: ndarray | list of str | dict}"""
    return File(path).root
