"""Minimal pure-Python reader for the NetCDF-4 / HDF5 files of the reference's test data.

Test infrastructure only (fixture generation, tests/golden/make_golden.py): this image has no
libhdf5 / netCDF4 / h5py, and every *.nc of the reference (forcings, restart files, gridded
outputs) is HDF5.  Implements just what those files use: superblock v0/v2, object headers v1/v2
with continuation blocks, old-style groups (v1 B-tree + local heap + SNOD), new-style groups
(link messages, compact or dense = fractal heap), dataspace / datatype (fixed, float, string) /
fill value / layout (compact, contiguous, chunked with v1 B-tree) / filter pipeline (deflate,
shuffle, fletcher32) messages and attributes (compact or dense).

    f = H5File(path); f.keys(); f["L1_fSealed"].read(); f["pre"].attrs
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class _Buf:
    def __init__(self, data, pos=0):
        self.d, self.p = data, pos

    def u(self, n):
        v = int.from_bytes(self.d[self.p:self.p + n], "little")
        self.p += n
        return v

    def raw(self, n):
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def skip(self, n):
        self.p += n

    def align(self, base, a=8):
        r = (self.p - base) % a
        if r:
            self.p += a - r


class Dataset:
    def __init__(self, f, name, msgs):
        self.f, self.name, self.msgs = f, name, msgs
        self.shape = ()
        self.dtype = None
        self.layout = None
        self.filters = []
        self.fill = None
        self.attrs = {}
        self._parse()

    def _parse(self):
        f = self.f
        for t, body in self.msgs:
            if t == 0x01:
                self.shape = f._dataspace(body)
            elif t == 0x03:
                self.dtype = f._datatype(_Buf(body))
            elif t == 0x08:
                self.layout = f._layout(body)
            elif t == 0x0B:
                self.filters = f._filters(body)
            elif t == 0x05:
                self.fill_raw = f._fillvalue(body)
            elif t == 0x0C:
                k, v = f._attribute(body)
                self.attrs[k] = v
            elif t == 0x15:
                self.attrs.update(f._dense_attrs(body))

    @property
    def is_dataset(self):
        return self.layout is not None and self.dtype is not None

    def read(self):
        f, dt = self.f, self.dtype
        if isinstance(dt, tuple):
            raise H5Error("%s: unsupported datatype %r" % (self.name, dt))
        n = int(np.prod(self.shape)) if self.shape else 1
        kind = self.layout[0]
        if kind == "compact":
            return np.frombuffer(self.layout[1], dtype=dt, count=n).reshape(self.shape).copy()
        if kind == "contiguous":
            addr, size = self.layout[1], self.layout[2]
            if addr == UNDEF:
                return self._filled(self.shape)
            return np.frombuffer(f.d, dtype=dt, count=n, offset=addr).reshape(self.shape).copy()
        addr, cdims = self.layout[1], self.layout[2]
        out = self._filled(self.shape)
        if addr == UNDEF:
            return out
        rank = len(self.shape)
        for offs, caddr, csize, mask in f._chunks(addr, rank):
            raw = f.d[caddr:caddr + csize]
            for i, (fid, cd) in reversed(list(enumerate(self.filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = dt.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = len(a) // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + bytes(a[m * es:])
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Error("filter %d not supported" % fid)
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sl_o, sl_c = [], []
            for o, c, s in zip(offs, cdims, self.shape):
                e = min(o + c, s)
                sl_o.append(slice(o, e))
                sl_c.append(slice(0, e - o))
            out[tuple(sl_o)] = chunk[tuple(sl_c)]
        return out

    def _filled(self, shape):
        out = np.zeros(shape, dtype=self.dtype)
        raw = getattr(self, "fill_raw", None)
        if raw:
            out[...] = np.frombuffer(raw[:self.dtype.itemsize], dtype=self.dtype)[0]
        return out


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.d = fh.read()
        if self.d[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file: " + path)
        ver = self.d[8]
        if ver in (0, 1):
            self.O, self.L = self.d[13], self.d[14]
            b = _Buf(self.d, 24 if ver == 0 else 28)
            b.skip(4 * self.O)          # base, free space, eof, driver
            b.skip(self.O)              # link name offset
            root = b.u(self.O)
        elif ver in (2, 3):
            self.O, self.L = self.d[9], self.d[10]
            b = _Buf(self.d, 12)
            b.skip(3 * self.O)
            root = b.u(self.O)
        else:
            raise H5Error("superblock version %d" % ver)
        self.root = self._object(root, "/")
        self._links = self._group_links(self.root.msgs)
        self._cache = {}

    # ---- public ------------------------------------------------------------------------
    def keys(self):
        return list(self._links)

    def __contains__(self, k):
        return k in self._links

    def __getitem__(self, k):
        if k not in self._cache:
            self._cache[k] = self._object(self._links[k], k)
        return self._cache[k]

    @property
    def attrs(self):
        return self.root.attrs

    # ---- object headers ----------------------------------------------------------------
    def _object(self, addr, name):
        return Dataset(self, name, self._messages(addr))

    def _messages(self, addr):
        d = self.d
        msgs = []
        if d[addr:addr + 4] == b"OHDR":
            b = _Buf(d, addr + 4)
            if b.u(1) != 2:
                raise H5Error("OHDR version")
            flags = b.u(1)
            if flags & 0x20:
                b.skip(16)
            if flags & 0x10:
                b.skip(4)
            size = b.u(1 << (flags & 3))
            blocks = [(b.p, b.p + size)]
            track = bool(flags & 0x04)
            while blocks:
                p, e = blocks.pop(0)
                b = _Buf(d, p)
                while b.p + 4 <= e:
                    t = b.u(1)
                    sz = b.u(2)
                    b.u(1)
                    if track:
                        b.skip(2)
                    body = b.raw(sz)
                    if t == 0x10:
                        cb = _Buf(body)
                        ca, cl = cb.u(self.O), cb.u(self.L)
                        if d[ca:ca + 4] != b"OCHK":
                            raise H5Error("OCHK expected")
                        blocks.append((ca + 4, ca + cl - 4))
                    elif t != 0:
                        msgs.append((t, body))
            return msgs
        # version 1
        b = _Buf(d, addr)
        if b.u(1) != 1:
            raise H5Error("object header version at %d" % addr)
        b.skip(1)
        nmsg = b.u(2)
        b.skip(4)
        size = b.u(4)
        b.align(addr, 8)
        blocks = [(b.p, b.p + size)]
        while blocks and nmsg > 0:
            p, e = blocks.pop(0)
            b = _Buf(d, p)
            while b.p + 8 <= e and nmsg > 0:
                t = b.u(2)
                sz = b.u(2)
                b.skip(4)
                body = b.raw(sz)
                nmsg -= 1
                if t == 0x10:
                    cb = _Buf(body)
                    ca, cl = cb.u(self.O), cb.u(self.L)
                    blocks.append((ca, ca + cl))
                elif t != 0:
                    msgs.append((t, body))
        return msgs

    # ---- groups ------------------------------------------------------------------------
    def _group_links(self, msgs):
        links = {}
        for t, body in msgs:
            if t == 0x11:
                b = _Buf(body)
                links.update(self._old_group(b.u(self.O), b.u(self.O)))
            elif t == 0x06:
                k, a = self._link(_Buf(body))
                if a is not None:
                    links[k] = a
            elif t == 0x02:
                b = _Buf(body)
                b.skip(1)
                fl = b.u(1)
                if fl & 1:
                    b.skip(8)
                heap = b.u(self.O)
                if heap != UNDEF:
                    for obj in self._fractal_objects(heap, lambda x: x[0] == 1):
                        k, a = self._link(_Buf(obj))
                        if a is not None:
                            links[k] = a
        return links

    def _link(self, b):
        if b.u(1) != 1:
            raise H5Error("link version")
        fl = b.u(1)
        ltype = b.u(1) if fl & 0x08 else 0
        if fl & 0x04:
            b.skip(8)
        if fl & 0x10:
            b.skip(1)
        n = b.u(1 << (fl & 3))
        name = b.raw(n).decode()
        if ltype != 0:
            return name, None
        return name, b.u(self.O)

    def _link_len(self, data, p):
        """length of the link message starting at data[p] (for the sequential heap scan)"""
        b = _Buf(data, p)
        b.u(1)
        fl = b.u(1)
        ltype = b.u(1) if fl & 0x08 else 0
        if fl & 0x04:
            b.skip(8)
        if fl & 0x10:
            b.skip(1)
        n = b.u(1 << (fl & 3))
        b.skip(n)
        if ltype == 0:
            b.skip(self.O)
        elif ltype == 1:
            b.skip(b.u(2))
        else:
            b.skip(b.u(2))
        return b.p - p

    def _old_group(self, btree, heap):
        d = self.d
        if d[heap:heap + 4] != b"HEAP":
            raise H5Error("HEAP expected")
        hb = _Buf(d, heap + 8)
        hb.skip(2 * self.L)
        hdata = hb.u(self.O)
        links = {}

        def walk(addr):
            if d[addr:addr + 4] == b"TREE":
                b = _Buf(d, addr + 4)
                b.u(1)
                b.u(1)
                n = b.u(2)
                b.skip(2 * self.O)
                for _ in range(n):
                    b.skip(self.L)
                    walk(b.u(self.O))
            elif d[addr:addr + 4] == b"SNOD":
                b = _Buf(d, addr + 6)
                n = b.u(2)
                for _ in range(n):
                    no = b.u(self.O)
                    oa = b.u(self.O)
                    b.skip(24)
                    e = d.index(b"\0", hdata + no)
                    links[d[hdata + no:e].decode()] = oa
            else:
                raise H5Error("TREE/SNOD expected")
        walk(btree)
        return links

    # ---- fractal heap (dense links / attributes): sequential scan of the direct blocks --
    def _fractal_objects(self, addr, is_start):
        d = self.d
        if d[addr:addr + 4] != b"FRHP":
            raise H5Error("FRHP expected")
        b = _Buf(d, addr + 5)
        b.skip(2)                       # heap id length
        filt_len = b.u(2)
        flags = b.u(1)
        b.skip(4)                       # max managed object size
        b.skip(self.L + self.O)         # next huge id, huge btree
        b.skip(self.L + self.O)         # free space, free space manager
        b.skip(4 * self.L)              # managed space, allocated, iterator offset, n managed
        n_managed = _Buf(d, b.p - self.L).u(self.L)
        b.skip(4 * self.L)              # huge size/n, tiny size/n
        width = b.u(2)
        start_size = b.u(self.L)
        max_direct = b.u(self.L)
        max_heap_bits = b.u(2)
        b.skip(2)                       # start rows
        root = b.u(self.O)
        cur_rows = b.u(2)
        if filt_len:
            raise H5Error("filtered fractal heap")
        off_bytes = (max_heap_bits + 7) // 8
        checksum = bool(flags & 2)
        objs = []

        def direct(a, size):
            if d[a:a + 4] != b"FHDB":
                raise H5Error("FHDB expected")
            p = a + 5 + self.O + off_bytes + (4 if checksum else 0)
            e = a + size
            while p < e and is_start(d[p:p + 2]):
                n = self._scan_len(d, p)
                objs.append(d[p:p + n])
                p += n

        def row_size(r):
            return start_size if r < 2 else start_size << (r - 1)

        def indirect(a, nrows):
            if d[a:a + 4] != b"FHIB":
                raise H5Error("FHIB expected")
            bb = _Buf(d, a + 5 + self.O + off_bytes)
            max_direct_rows = 2
            s = start_size
            while s < max_direct:
                s <<= 1
                max_direct_rows += 1
            for r in range(nrows):
                for _ in range(width):
                    ca = bb.u(self.O)
                    if r < max_direct_rows:
                        if ca != UNDEF:
                            direct(ca, row_size(r))
                    elif ca != UNDEF:
                        raise H5Error("nested indirect fractal-heap blocks not supported")

        if root == UNDEF:
            return objs
        if cur_rows == 0:
            direct(root, start_size)
        else:
            indirect(root, cur_rows)
        if len(objs) != n_managed:
            raise H5Error("fractal heap scan found %d of %d objects" % (len(objs), n_managed))
        return objs

    def _scan_len(self, d, p):
        return self._scanner(d, p)

    # set per call by _dense_attrs / _group_links
    def _scanner(self, d, p):
        return self._link_len(d, p)

    # ---- messages ----------------------------------------------------------------------
    def _dataspace(self, body):
        b = _Buf(body)
        ver = b.u(1)
        rank = b.u(1)
        fl = b.u(1)
        if ver == 1:
            b.skip(5)
        else:
            b.skip(1)
        return tuple(b.u(self.L) for _ in range(rank))

    def _datatype(self, b):
        cv = b.u(1)
        cls, bits0 = cv & 0x0F, b.u(1)
        b.skip(2)
        size = b.u(4)
        if cls == 0:
            b.skip(4)
            return np.dtype(("<" if not bits0 & 1 else ">") + ("i" if bits0 & 8 else "u") + str(size))
        if cls == 1:
            b.skip(12)
            return np.dtype(("<" if not bits0 & 1 else ">") + "f" + str(size))
        if cls == 3:
            return np.dtype("S%d" % size)
        return ("class", cls, size)

    def _fillvalue(self, body):
        b = _Buf(body)
        ver = b.u(1)
        if ver in (1, 2):
            b.skip(2)
            defined = b.u(1)
            if ver == 1 or defined:
                n = b.u(4)
                return b.raw(n)
            return None
        fl = b.u(1)
        if fl & 0x20:
            n = b.u(4)
            return b.raw(n)
        return None

    def _layout(self, body):
        b = _Buf(body)
        ver = b.u(1)
        if ver != 3:
            raise H5Error("data layout version %d" % ver)
        cls = b.u(1)
        if cls == 0:
            n = b.u(2)
            return ("compact", b.raw(n))
        if cls == 1:
            return ("contiguous", b.u(self.O), b.u(self.L))
        nd = b.u(1)
        addr = b.u(self.O)
        dims = [b.u(4) for _ in range(nd)]
        return ("chunked", addr, tuple(dims[:-1]))

    def _filters(self, body):
        b = _Buf(body)
        ver = b.u(1)
        n = b.u(1)
        out = []
        if ver == 1:
            b.skip(6)
        for _ in range(n):
            fid = b.u(2)
            nl = b.u(2) if (ver == 1 or fid >= 256) else 0
            b.skip(2)
            ncd = b.u(2)
            if ver == 1:
                b.skip((nl + 7) // 8 * 8)
            else:
                b.skip(nl)
            cd = [b.u(4) for _ in range(ncd)]
            if ver == 1 and ncd % 2:
                b.skip(4)
            out.append((fid, cd))
        return out

    def _chunks(self, addr, rank):
        d = self.d
        if d[addr:addr + 4] != b"TREE":
            raise H5Error("chunk TREE expected")
        b = _Buf(d, addr + 4)
        if b.u(1) != 1:
            raise H5Error("chunk btree node type")
        level = b.u(1)
        n = b.u(2)
        b.skip(2 * self.O)
        for _ in range(n):
            csize = b.u(4)
            mask = b.u(4)
            offs = [b.u(8) for _ in range(rank + 1)][:rank]
            child = b.u(self.O)
            if level == 0:
                yield offs, child, csize, mask
            else:
                yield from self._chunks(child, rank)

    def _attribute(self, body):
        b = _Buf(body)
        ver = b.u(1)
        b.u(1)
        ns, ts, ss = b.u(2), b.u(2), b.u(2)
        if ver == 3:
            b.skip(1)
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        name = b.raw(ns).split(b"\0")[0].decode()
        b.skip(pad(ns) - ns)
        tb = b.raw(ts)
        b.skip(pad(ts) - ts)
        sb = b.raw(ss)
        b.skip(pad(ss) - ss)
        dt = self._datatype(_Buf(tb))
        shape = self._dataspace(sb) if ss else ()
        if isinstance(dt, tuple):
            return name, None
        n = int(np.prod(shape)) if shape else 1
        try:
            v = np.frombuffer(body, dtype=dt, count=n, offset=b.p)
        except ValueError:
            return name, None
        if dt.kind == "S":
            v = v[0].split(b"\0")[0].decode(errors="replace") if n == 1 else [x.decode() for x in v]
        elif not shape:
            v = v[0]
        return name, v

    def _attr_len(self, d, p):
        b = _Buf(d, p)
        ver = b.u(1)
        b.u(1)
        ns, ts, ss = b.u(2), b.u(2), b.u(2)
        if ver == 3:
            b.skip(1)
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        tb = d[b.p + pad(ns): b.p + pad(ns) + ts]
        sb = d[b.p + pad(ns) + pad(ts): b.p + pad(ns) + pad(ts) + ss]
        dt = self._datatype(_Buf(tb))
        shape = self._dataspace(sb) if ss else ()
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, tuple):
            es = dt[2] if dt[1] != 9 else 4 + self.O + 4
        else:
            es = dt.itemsize
        return (b.p - p) + pad(ns) + pad(ts) + pad(ss) + n * es

    def _dense_attrs(self, body):
        b = _Buf(body)
        b.skip(1)
        fl = b.u(1)
        if fl & 1:
            b.skip(2)
        heap = b.u(self.O)
        out = {}
        if heap == UNDEF:
            return out
        self._scanner = self._attr_len
        try:
            for obj in self._fractal_objects(heap, lambda x: len(x) == 2 and x[0] in (1, 2, 3) and x[1] in (0, 1, 2, 3)):
                k, v = self._attribute(obj)
                out[k] = v
        finally:
            self._scanner = self._link_len
        return out
