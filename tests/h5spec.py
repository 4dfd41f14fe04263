"""An independent, strict pure-Python parser of the HDF5 structures the reference's files use (superblock v0, v1 object headers,
symbol-table groups: B-tree v1 + local heap + SNOD, contiguous layout), written from the HDF5 File Format Specification and NOT
sharing a line with host/h5lite.cpp.  Test infrastructure: tests/test_host_h5.py pins it on a libhdf5-written file and then uses it
as a second opinion on what h5lite's Writer emits.  Every `_ck` is a condition libhdf5 enforces when it opens / reads the file."""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class FormatError(Exception):
    pass


def _ck(cond, what):
    if not cond:
        raise FormatError(what)


class File:
    def __init__(self, path, strict_eof=True):  # strict_eof: nothing may follow the end-of-file address
        self.b = open(path, "rb").read()
        b = self.b
        sb = 0
        while b[sb:sb + 8] != SIG:  # the superblock sits at 0, 512, 1024, ... (II.A)
            sb = 512 if sb == 0 else sb * 2
            _ck(sb + 8 <= len(b), "no superblock signature")
        self.user_block = sb
        ver, fs_ver, ste_ver, r0, shm_ver, self.O, self.L, r1 = struct.unpack_from("<8B", b, sb + 8)
        _ck(ver == 0 and fs_ver == 0 and ste_ver == 0 and shm_ver == 0 and r0 == 0 and r1 == 0, "superblock v0 version / reserved bytes")
        _ck(self.O == 8 and self.L == 8, "8-byte offsets and lengths")
        self.leaf_k, self.int_k, flags = struct.unpack_from("<HHI", b, sb + 16)
        _ck(self.leaf_k > 0 and self.int_k > 0, "group K values")
        base, fsinfo, eof, drv = struct.unpack_from("<4Q", b, sb + 24)
        _ck(fsinfo == UNDEF and drv == UNDEF, "no free-space / driver info blocks")
        _ck(base == sb, "base address == position of the superblock (= size of the user block)")
        self.base = base
        # the stored end-of-file address is absolute (libhdf5 adds the base address to the file's length before comparing)
        _ck(eof == len(b) if strict_eof else eof <= len(b), f"end-of-file address {eof} vs file size {len(b)}")
        self.eof = eof
        name_off, oh, cache, rsv = struct.unpack_from("<QQII", b, sb + 56)
        _ck(name_off == 0 and rsv == 0, "root entry")
        self.root_oh = oh
        root = self.object_header(oh)
        stab = [m for m in root if m[0] == 0x0011]
        _ck(len(stab) == 1, "root group has one symbol-table message")
        btree, heap = struct.unpack_from("<QQ", stab[0][2], 0)
        if cache == 1:  # the scratch pad caches the same two addresses (H5G__stab_valid)
            _ck(struct.unpack_from("<QQ", b, sb + 80) == (btree, heap), "root entry scratch pad == symbol-table message")
        self.heap = self._heap(heap)
        self.links = {}
        self._group_node(btree, None, None, 0)

    # ------------------------------------------------------------------ primitives
    def at(self, addr, n):
        _ck(addr != UNDEF and self.base + addr + n <= self.eof, f"address {addr}+{n} outside the file")
        return self.b[self.base + addr:self.base + addr + n]

    def _heap(self, addr):
        h = self.at(addr, 32)
        _ck(h[:4] == b"HEAP" and h[4] == 0 and h[5:8] == b"\0\0\0", "local heap header")
        size, free, data = struct.unpack_from("<QQQ", h, 8)
        seg = self.at(data, size)
        _ck(size % 8 == 0, "heap segment is 8-byte aligned")
        seen = 0
        while free != 1:  # free list: (next, size) pairs inside the segment, 1 = end of list (H5HL_FREE_NULL)
            _ck(free % 8 == 0 and free + 16 <= size, "free block inside the heap segment")
            nxt, fsz = struct.unpack_from("<QQ", seg, free)
            _ck(fsz >= 16 and free + fsz <= size, "free block size")
            free = nxt
            seen += 1
            _ck(seen < 1000, "free list loop")
        return seg

    def name(self, off):
        _ck(off < len(self.heap), "name offset inside the heap")
        end = self.heap.index(b"\0", off)
        return self.heap[off:end].decode("ascii")

    def _group_node(self, addr, lo, hi, depth):
        h = self.at(addr, 24)
        _ck(h[:4] == b"TREE" and h[4] == 0, "group B-tree node")
        level, used = h[5], struct.unpack_from("<H", h, 6)[0]
        _ck(used <= 2 * self.int_k, "B-tree entries used <= 2K")
        body = self.at(addr + 24, (2 * self.int_k + 1) * 8 + 2 * self.int_k * 8)
        keys = [struct.unpack_from("<Q", body, 16 * i)[0] for i in range(used + 1)]
        kids = [struct.unpack_from("<Q", body, 16 * i + 8)[0] for i in range(used)]
        knames = [self.name(k) for k in keys]
        _ck(all(knames[i] < knames[i + 1] for i in range(used)) or used == 0, "B-tree keys strictly ascending")
        for i, c in enumerate(kids):
            if level:
                self._group_node(c, knames[i], knames[i + 1], depth + 1)
            else:
                self._snod(c, knames[i], knames[i + 1])

    def _snod(self, addr, lo, hi):
        h = self.at(addr, 8 + 2 * self.leaf_k * 40)
        _ck(h[:4] == b"SNOD" and h[4] == 1 and h[5] == 0, "symbol table node header")
        n = struct.unpack_from("<H", h, 6)[0]
        _ck(0 < n <= 2 * self.leaf_k, "symbols per node")
        prev = None
        for i in range(n):
            off, oh, cache, rsv = struct.unpack_from("<QQII", h, 8 + 40 * i)
            nm = self.name(off)
            _ck(rsv == 0 and cache in (0, 1, 2), "symbol table entry")
            _ck(prev is None or prev < nm, "SNOD entries sorted by name")
            _ck(lo < nm <= hi, f"name {nm!r} inside its B-tree key interval ({lo!r}, {hi!r}]")
            _ck(nm not in self.links, "duplicate link")
            self.links[nm] = oh
            prev = nm

    def object_header(self, addr):
        """-> [(type, flags, body)] of a version-1 object header, continuation blocks followed."""
        h = self.at(addr, 16)
        _ck(h[0] == 1 and h[1] == 0, "object header version 1")
        nmsg, refs, size = struct.unpack_from("<HII", h, 2)
        _ck(refs >= 1, "reference count")
        _ck(addr % 8 == 0 and size % 8 == 0, "object header alignment")
        blocks, out = [(addr + 16, size)], []
        while blocks:
            a, n = blocks.pop(0)
            blk, p = self.at(a, n), 0
            while p + 8 <= n:
                t, sz, fl = struct.unpack_from("<HHB", blk, p)
                _ck(blk[p + 5:p + 8] == b"\0\0\0", "message header reserved bytes")
                _ck(sz % 8 == 0 and p + 8 + sz <= n, f"message 0x{t:04x}: size {sz} aligned and inside its block")
                body = blk[p + 8:p + 8 + sz]
                if t == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                out.append((t, fl, body))
                p += 8 + sz
            _ck(p == n or n - p < 8, "messages fill the block")
        _ck(len(out) == nmsg, f"object header announces {nmsg} messages, holds {len(out)}")  # H5O: 'incorrect # of messages'
        return out

    # ------------------------------------------------------------------ datasets
    def dataset(self, name):
        msgs = {t: body for t, _, body in self.object_header(self.links[name])}
        _ck(0x0001 in msgs and 0x0003 in msgs and 0x0008 in msgs, "dataset has dataspace, datatype and layout messages")
        sp = msgs[0x0001]
        _ck(sp[0] in (1, 2), "dataspace version")
        rank, flags = sp[1], sp[2]
        p = 8 if sp[0] == 1 else 4
        if sp[0] == 1:
            _ck(sp[3:8] == b"\0" * 5, "dataspace v1 reserved bytes")
        dims = struct.unpack_from(f"<{rank}Q", sp, p)
        _ck(len(sp) >= p + 8 * rank * (2 if flags & 1 else 1), "dataspace message holds its dimensions")
        dt = self._dtype(msgs[0x0003])
        lay = msgs[0x0008]
        count = int(np.prod(dims, dtype=np.uint64)) if rank else 1
        if lay[0] == 3:
            _ck(lay[1] == 1, "data layout version 3: contiguous")
            addr, nbytes = struct.unpack_from("<QQ", lay, 2)
        else:  # versions 1 and 2 (HDF5 1.6): dimensionality (rank + 1: the element size comes last), class, 5 reserved bytes, address, 32-bit sizes
            _ck(lay[0] in (1, 2) and lay[2] == 1 and lay[1] == rank + 1 and lay[3:8] == b"\0" * 5, "data layout version 1/2: contiguous")
            addr = struct.unpack_from("<Q", lay, 8)[0]
            _ck(struct.unpack_from(f"<{rank + 1}I", lay, 16) == tuple(dims) + (dt.itemsize,), "layout dimensions == dataspace dimensions + element size")
            nbytes = count * dt.itemsize
        _ck(nbytes == count * dt.itemsize, "layout size == elements x element size")
        if 0x0005 in msgs:
            fv = msgs[0x0005]
            _ck(fv[0] in (1, 2, 3), "fill value version")
            if fv[0] == 2 and fv[3]:
                _ck(len(fv) >= 8, "fill value v2 with 'defined' carries a size field")
        if nbytes == 0:
            return np.zeros(dims, dt)
        return np.frombuffer(self.at(addr, nbytes), dt).reshape(dims)

    @staticmethod
    def _dtype(m):
        cls, ver = m[0] & 15, m[0] >> 4
        _ck(ver in (1, 2, 3), "datatype version")
        bits = m[1] | (m[2] << 8) | (m[3] << 16)
        size = struct.unpack_from("<I", m, 4)[0]
        order = ">" if bits & 1 else "<"
        off, prec = struct.unpack_from("<HH", m, 8)
        _ck(off == 0 and prec == 8 * size, "no bit padding")
        if cls == 0:
            _ck(bits & ~0x9 == 0 and size in (1, 2, 4, 8), "fixed-point class bits")
            return np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}")
        _ck(cls == 1, f"datatype class {cls}")
        eloc, esz, mloc, msz = m[12:16]
        bias = struct.unpack_from("<I", m, 16)[0]
        sign = (bits >> 8) & 255
        _ck((bits >> 4) & 3 == 2 and bits & 0x4E == 0, "IEEE: implied mantissa msb, zero padding")
        _ck((size, sign, eloc, esz, mloc, msz, bias) in ((4, 31, 23, 8, 0, 23, 127), (8, 63, 52, 11, 0, 52, 1023)), "IEEE binary32 / binary64 fields")
        return np.dtype(f"{order}f{size}")
