"""Dependency-free reader for the HDF5 subset the shipped Exodus meshes use (NetCDF-4 written through h5py):

superblock version 0, version-1 object headers (+ continuation blocks), old-style groups (symbol-table message ->
version-1 B-tree -> symbol nodes, names in a local heap), datasets with compact / contiguous / chunked layout (version-3
layout message, version-1 chunk B-tree), filter pipeline versions 1 / 2 with shuffle (2) and deflate (1), fixed-point,
floating-point and fixed-length string datatypes, simple dataspaces; scalar / 1-D attributes (version 1-3 messages).

This stands in for the netCDF/HDF5 libraries behind the reference's NetCDF_Reader (SOLVER/src/preloop/utilities/netcdf/
NetCDF_Reader.cpp), which are not in this image.  Anything outside the subset raises H5Error naming the feature.
Format reference: "HDF5 File Format Specification Version 2.0" (public), restated here.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class Dataset:
    def __init__(self, f, name, msgs):
        self.f, self.name = f, name
        self.shape, self.dtype, self.layout, self.filters, self.attrs_raw = None, None, None, [], []
        for t, d in msgs:
            if t == 0x01:
                self.shape = _dataspace(d)
            elif t == 0x03:
                self.dtype = _datatype(d)
            elif t == 0x08:
                self.layout = _layout(d)
            elif t == 0x0B:
                self.filters = _filters(d)
            elif t == 0x0C:
                self.attrs_raw.append(d)
        if self.shape is None or self.dtype is None or self.layout is None:
            raise H5Error("%s: not a dataset" % name)

    @property
    def attrs(self):
        out = {}
        for d in self.attrs_raw:
            try:
                k, v = _attribute(d)
                out[k] = v
            except H5Error:
                pass
        return out

    def read(self):
        dt, itemsize = self.dtype
        n = int(np.prod(self.shape)) if self.shape else 1
        buf = self.f.buf
        kind = self.layout[0]
        if kind == "compact":
            raw = self.layout[1]
        elif kind == "contiguous":
            addr, size = self.layout[1], self.layout[2]
            raw = b"\0" * (n * itemsize) if addr == UNDEF else buf[addr:addr + size]
        else:
            raw = self._read_chunked(n, itemsize)
        a = np.frombuffer(raw[:n * itemsize], dtype=dt)
        return a.reshape(self.shape) if self.shape else a.reshape(())

    def _read_chunked(self, n, itemsize):
        _, btree, cdims = self.layout          # cdims includes the trailing element size
        rank = len(self.shape)
        cshape = tuple(cdims[:rank])
        out = np.zeros(self.shape, dtype=np.dtype("V%d" % itemsize))
        if btree == UNDEF:
            return out.tobytes()
        for off, size, mask, addr in self.f._chunks(btree, rank):
            raw = self.f.buf[addr:addr + size]
            for k in reversed(range(len(self.filters))):
                fid, cd = self.filters[k]
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else itemsize
                    m = len(raw) // es
                    raw = np.frombuffer(raw[:m * es], dtype=np.uint8).reshape(es, m).T.tobytes() + raw[m * es:]
                elif fid == 3:
                    raw = raw[:-4]              # fletcher32 checksum trailer
                else:
                    raise H5Error("%s: filter id %d" % (self.name, fid))
            c = np.frombuffer(raw, dtype=out.dtype, count=int(np.prod(cshape))).reshape(cshape)
            sl_out, sl_in = [], []
            for d in range(rank):
                lo = off[d]
                hi = min(lo + cshape[d], self.shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = c[tuple(sl_in)]
        return out.tobytes()


class File:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        if b[8] != 0:
            raise H5Error("superblock version %d" % b[8])
        if b[13] != 8 or b[14] != 8:
            raise H5Error("offset / length size %d / %d" % (b[13], b[14]))
        self.base = struct.unpack_from("<Q", b, 24)[0]
        # root group symbol table entry at 56: link name offset, object header address, cache type, reserved, scratch
        _, oh, cache = struct.unpack_from("<QQI", b, 56)
        self.root = oh
        self._index = None

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        b = self.buf
        ver, _, nmsg, _, hsz = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error("object header version %d" % ver)
        out, blocks = [], [(addr + 16, addr + 16 + hsz)]
        while blocks:
            q, end = blocks.pop(0)
            while q + 8 <= end and len(out) < nmsg:
                t, sz, _ = struct.unpack_from("<HHB", b, q)
                d = b[q + 8:q + 8 + sz]
                if t == 0x10:
                    o, l = struct.unpack_from("<QQ", d)
                    blocks.append((o, o + l))
                    out.append((t, d))
                else:
                    out.append((t, d))
                q += 8 + sz
        return out

    # ------------------------------------------------------------------ groups (old style)
    def _children(self, addr):
        for t, d in self._messages(addr):
            if t == 0x11:
                bt, heap = struct.unpack_from("<QQ", d)
                return self._walk_group(bt, heap)
            if t in (0x02, 0x06):
                raise H5Error("new-style group (link messages)")
        return None

    def _heap_name(self, heap, off):
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error("local heap signature")
        data = struct.unpack_from("<Q", b, heap + 24)[0]
        e = b.index(b"\0", data + off)
        return b[data + off:e].decode()

    def _walk_group(self, bt, heap):
        b = self.buf
        out = {}
        if b[bt:bt + 4] != b"TREE":
            raise H5Error("B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, bt + 4)
        if ntype != 0:
            raise H5Error("group B-tree node type %d" % ntype)
        p = bt + 24
        for k in range(used):
            child = struct.unpack_from("<Q", b, p + 8)[0]
            p += 16
            if level > 0:
                out.update(self._walk_group(child, heap))
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5Error("symbol node signature")
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                q = child + 8
                for _ in range(nsym):
                    lno, oh = struct.unpack_from("<QQ", b, q)
                    out[self._heap_name(heap, lno)] = oh
                    q += 40
        return out

    def index(self):
        """{path: object header address} of every dataset / group below the root"""
        if self._index is None:
            idx = {}

            def rec(prefix, addr):
                ch = self._children(addr)
                if ch is None:
                    return
                for name, oh in ch.items():
                    idx[prefix + name] = oh
                    rec(prefix + name + "/", oh)

            rec("", self.root)
            self._index = idx
        return self._index

    def keys(self):
        return sorted(self.index())

    def __contains__(self, name):
        return name in self.index()

    def __getitem__(self, name):
        idx = self.index()
        if name not in idx:
            raise KeyError(name)
        return Dataset(self, name, self._messages(idx[name]))

    def root_attrs(self):
        out = {}
        for t, d in self._messages(self.root):
            if t == 0x0C:
                try:
                    k, v = _attribute(d)
                    out[k] = v
                except H5Error:
                    pass
        return out

    # ------------------------------------------------------------------ chunk B-tree (version 1, node type 1)
    def _chunks(self, bt, rank):
        b = self.buf
        if b[bt:bt + 4] != b"TREE":
            raise H5Error("chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, bt + 4)
        if ntype != 1:
            raise H5Error("chunk B-tree node type %d" % ntype)
        ksz = 8 + 8 * (rank + 1)
        p = bt + 24
        for k in range(used):
            size, mask = struct.unpack_from("<II", b, p)
            off = struct.unpack_from("<%dQ" % (rank + 1), b, p + 8)
            child = struct.unpack_from("<Q", b, p + ksz)[0]
            p += ksz + 8
            if level > 0:
                yield from self._chunks(child, rank)
            else:
                yield off[:rank], size, mask, child


# ---------------------------------------------------------------------- message decoders
def _dataspace(d):
    ver, rank, flags = struct.unpack_from("<BBB", d)
    if ver == 1:
        p = 8
    elif ver == 2:
        if d[3] == 2:
            return ()          # null dataspace
        p = 4
    else:
        raise H5Error("dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, d, p)) if rank else ()


def _datatype(d):
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", d)
    cls = cv & 0x0F
    be = ">" if (b0 & 1) else "<"
    if cls == 0:
        return np.dtype("%s%s%d" % (be, "i" if (b0 & 8) else "u", size)), size
    if cls == 1:
        return np.dtype("%sf%d" % (be, size)), size
    if cls == 3:
        return np.dtype("S%d" % size), size
    raise H5Error("datatype class %d" % cls)


def _layout(d):
    ver = d[0]
    if ver != 3:
        raise H5Error("layout message version %d" % ver)
    cls = d[1]
    if cls == 0:
        sz = struct.unpack_from("<H", d, 2)[0]
        return ("compact", d[4:4 + sz])
    if cls == 1:
        addr, size = struct.unpack_from("<QQ", d, 2)
        return ("contiguous", addr, size)
    if cls == 2:
        nd = d[2]
        addr = struct.unpack_from("<Q", d, 3)[0]
        dims = struct.unpack_from("<%dI" % nd, d, 11)
        return ("chunked", addr, dims)
    raise H5Error("layout class %d" % cls)


def _filters(d):
    ver, nf = d[0], d[1]
    out = []
    p = 8 if ver == 1 else 2
    for _ in range(nf):
        if ver == 1:
            fid, nlen, _, ncd = struct.unpack_from("<HHHH", d, p)
            p += 8
            p += (nlen + 7) // 8 * 8
        else:
            fid = struct.unpack_from("<H", d, p)[0]
            p += 2
            nlen = 0
            if fid >= 256:
                nlen = struct.unpack_from("<H", d, p)[0]
                p += 2
            _, ncd = struct.unpack_from("<HH", d, p)
            p += 4 + nlen
        cd = struct.unpack_from("<%dI" % ncd, d, p)
        p += 4 * ncd
        if ver == 1 and ncd % 2:
            p += 4
        out.append((fid, cd))
    return out


def _attribute(d):
    ver = d[0]
    if ver == 1:
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        p = 8
        pad = lambda n: (n + 7) // 8 * 8
    elif ver in (2, 3):
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        p = 8 + (1 if ver == 3 else 0)
        pad = lambda n: n
    else:
        raise H5Error("attribute version %d" % ver)
    name = d[p:p + nsz].split(b"\0")[0].decode()
    p += pad(nsz)
    dt, isz = _datatype(d[p:p + tsz])
    p += pad(tsz)
    shape = _dataspace(d[p:p + ssz])
    p += pad(ssz)
    n = int(np.prod(shape)) if shape else 1
    v = np.frombuffer(d[p:p + n * isz], dtype=dt)
    if dt.kind == "S":
        v = v[0].split(b"\0")[0].decode() if n == 1 else [x.split(b"\0")[0].decode() for x in v]
    elif n == 1:
        v = v[0].item()
    return name, v
