"""Independent, specification-driven walker of the HDF5 files the host driver writes (TEST
INFRASTRUCTURE).  It shares no code with the two readers of the repo (magnetizer_b200/hdf5_min.py,
host/mag_hdf5_reader.hpp) and checks, structure by structure, the invariants of the HDF5 File
Format Specification (version 0 superblock, version 1 object headers / B-trees / local heaps) that
libhdf5 relies on when it opens a file:

  superblock   signature, versions, 8-byte offsets/lengths, end-of-file address <= file size (== after a close),
               group leaf / internal K, root symbol-table entry with cached B-tree + heap addresses
  object hdr   version 1, message count and total size consistent, every message 8-byte aligned
  group        B-tree node type 0, entries <= 2K, node fits the file at its full size; local heap
               signature, data segment inside the file, free list well formed; symbol-table node
               version 1, entries <= 2*leafK, names inside the heap, sorted (strcmp order), keys of the
               B-tree = offsets of names that bound the children
  dataset      dataspace v1, datatype, fill value v2 (size == element size when defined), layout v3;
               contiguous: storage inside the file; chunked: dimensionality == rank+1, last chunk dim ==
               element size, B-tree node type 1, entries <= 2K (K = 32), levels consistent, keys strictly
               increasing (lexicographic offsets), left/right sibling addresses consistent on every level,
               right-most key beyond the last chunk, every chunk inside the file and inside the dataspace
               grid, aligned to the chunk shape, chunk byte size == product(chunk dims) * element size
               after the filter pipeline (deflate: zlib stream that inflates to exactly that size)
  attribute    version 1 message: name, datatype, dataspace sizes consistent

`walk(path)` returns {"/Group/dataset": Dataset} with the data decoded by this module.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
CHUNK_K = 32


class SpecError(AssertionError):
    pass


def need(cond, msg):
    if not cond:
        raise SpecError(msg)


class Dataset:
    def __init__(self):
        self.shape = None
        self.dtype = None
        self.fill = None
        self.layout = None
        self.chunk = None
        self.filters = []
        self.attrs = {}
        self.nchunks = 0
        self.btree_levels = 0
        self.data = None


class Walker:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        self.n = len(self.buf)
        self.objects = {}

    def u(self, fmt, off):
        need(off + struct.calcsize(fmt) <= self.n, f"read past end of file at {off}")
        return struct.unpack_from("<" + fmt, self.buf, off)

    # ---- superblock
    def superblock(self):
        b = self.buf
        need(b[:8] == b"\x89HDF\r\n\x1a\n", "signature")
        sb_ver, fs_ver, root_ver, _, sh_ver, so, sl, _ = b[8:16]
        need((sb_ver, fs_ver, root_ver, sh_ver) == (0, 0, 0, 0), "superblock versions")
        need((so, sl) == (8, 8), "offset/length sizes")
        self.leaf_k, self.int_k, flags = self.u("HHI", 16)
        need(self.leaf_k > 0 and self.int_k > 0 and flags == 0, "K values / consistency flags")
        base, free, eof, drv = self.u("QQQQ", 24)
        need(base == 0 and free == UNDEF and drv == UNDEF, "base / free-space / driver addresses")
        # libhdf5 refuses a file that is SHORTER than its end-of-file address (truncated); bytes beyond it --
        # what a killed run appended after its last flush -- are ignored
        need(eof <= self.n, f"end-of-file address {eof} beyond the file size {self.n}")
        self.trailing = self.n - eof
        name_off, ohdr, cache, _, btree, heap = self.u("QQIIQQ", 56)
        need(name_off == 0 and cache == 1, "root symbol-table entry must cache its B-tree and heap")
        return ohdr, btree, heap

    # ---- object header v1 -> list of (type, flags, body offset, size)
    def messages(self, ohdr):
        need(ohdr % 8 == 0, "object header alignment")
        ver, _, nmsg, refs, size = self.u("BBHII", ohdr)
        need(ver == 1 and refs >= 1, "object header version / reference count")
        off, end, out = ohdr + 16, ohdr + 16 + size, []
        need(end <= self.n, "object header inside file")
        while off + 8 <= end and len(out) < nmsg:
            mtype, msize, mflags = self.u("HHB", off)
            need(msize % 8 == 0, f"message 0x{mtype:x} size not a multiple of 8")
            need(mtype != 0x10, "continuation blocks are not expected in these files")
            out.append((mtype, mflags, off + 8, msize))
            off += 8 + msize
        need(len(out) == nmsg and off == end, "message count / header size mismatch")
        return out

    # ---- group
    def group(self, ohdr, btree_cached=None, heap_cached=None):
        msgs = self.messages(ohdr)
        stab = [m for m in msgs if m[0] == 0x11]
        need(len(stab) == 1, "group needs exactly one symbol-table message")
        btree, heap = self.u("QQ", stab[0][2])
        if btree_cached is not None:
            need((btree, heap) == (btree_cached, heap_cached), "cached B-tree/heap addresses differ from the symbol-table message")
        need(self.buf[heap:heap + 4] == b"HEAP" and self.buf[heap + 4] == 0, "local heap signature/version")
        seg_size, free_head, seg = self.u("QQQ", heap + 8)
        need(seg + seg_size <= self.n, "heap data segment inside file")
        if free_head != 1:   # H5HL_FREE_NULL
            need(free_head + 16 <= seg_size, "heap free-list head inside segment")
            nxt, fsz = self.u("QQ", seg + free_head)
            need(nxt == 1 and free_head + fsz <= seg_size, "heap free block")

        def name_at(o):
            need(o < seg_size, "name offset inside heap")
            e = self.buf.index(b"\0", seg + o)
            need(e < seg + seg_size, "name terminated inside heap")
            return self.buf[seg + o:e].decode()

        entries = []

        def node(addr, level_expected=None):
            need(addr % 8 == 0 and self.buf[addr:addr + 4] == b"TREE", "group B-tree node signature")
            ntype, level, nent = self.u("BBH", addr + 4)
            need(ntype == 0, "group B-tree node type")
            need(nent <= 2 * self.int_k, "B-tree node over capacity")
            need(addr + 24 + (2 * self.int_k + 1) * 8 + 2 * self.int_k * 8 <= self.n, "B-tree node fits the file at full size")
            if level_expected is not None:
                need(level == level_expected, "B-tree levels")
            keys = [self.u("Q", addr + 24 + 16 * i)[0] for i in range(nent + 1)]
            kids = [self.u("Q", addr + 24 + 16 * i + 8)[0] for i in range(nent)]
            knames = [name_at(k) for k in keys]
            need(all(knames[i] <= knames[i + 1] for i in range(nent)), "B-tree keys not sorted")
            for i, c in enumerate(kids):
                if level > 0:
                    node(c, level - 1)
                    continue
                need(self.buf[c:c + 4] == b"SNOD" and self.buf[c + 4] == 1, "symbol-table node signature/version")
                nsym = self.u("H", c + 6)[0]
                need(0 < nsym <= 2 * self.leaf_k, "symbol-table node entry count")
                need(c + 8 + 2 * self.leaf_k * 40 <= self.n, "symbol-table node fits the file at full size")
                names = []
                for s in range(nsym):
                    e = c + 8 + 40 * s
                    no, oh, cache = self.u("QQI", e)
                    nm = name_at(no)
                    names.append(nm)
                    scratch = self.u("QQ", e + 24) if cache == 1 else None
                    entries.append((nm, oh, scratch))
                need(names == sorted(names), "symbol-table entries not sorted")
                need(knames[i] < names[0], "left key must sort before the child's first name")
                need(knames[i + 1] == names[-1], "right key must be the child's last name")

        node(btree)
        attrs = self.attributes(msgs)
        return entries, attrs

    # ---- attributes (fixed-length strings / numbers)
    def attributes(self, msgs):
        out = {}
        for mtype, _, body, size in msgs:
            if mtype != 0x0C:
                continue
            ver, _, nlen, dtlen, dslen = self.u("BBHHH", body)
            need(ver == 1, "attribute message version")
            p = body + 8
            name = self.buf[p:p + nlen].split(b"\0")[0].decode()
            need(len(name) + 1 == nlen, "attribute name size")
            p += (nlen + 7) & ~7
            cls = self.buf[p] & 0x0F
            dsz = self.u("I", p + 4)[0]
            p += (dtlen + 7) & ~7
            dver, rank = self.buf[p], self.buf[p + 1]
            need(dver == 1, "attribute dataspace version")
            dims = [self.u("Q", p + 8 + 8 * i)[0] for i in range(rank)]
            p += (dslen + 7) & ~7
            nelem = int(np.prod(dims)) if dims else 1
            need(p + nelem * dsz <= body + size, "attribute data inside the message")
            raw = self.buf[p:p + nelem * dsz]
            out[name] = raw.decode(errors="replace").rstrip(" \0") if cls == 3 else raw
        return out

    # ---- dataset
    def dataset(self, ohdr):
        d = Dataset()
        msgs = self.messages(ohdr)
        types = [m[0] for m in msgs]
        for req in (0x01, 0x03, 0x05, 0x08):
            need(types.count(req) == 1, f"dataset needs exactly one message 0x{req:x}")
        layout_body = None
        for mtype, mflags, body, size in msgs:
            if mtype == 0x01:
                ver, rank, flags = self.u("BBB", body)
                need(ver == 1 and flags == 0, "dataspace version/flags")
                d.shape = tuple(self.u("Q", body + 8 + 8 * i)[0] for i in range(rank))
            elif mtype == 0x03:
                cv, b0 = self.buf[body], self.buf[body + 1]
                cls, ver = cv & 0x0F, cv >> 4
                need(ver == 1, "datatype version")
                sz = self.u("I", body + 4)[0]
                if cls == 1:
                    need(sz == 8 and not (b0 & 1), "only little-endian IEEE doubles expected")
                    off, prec, eloc, esz, mloc, msz, bias = self.u("HHBBBBI", body + 8)
                    need((off, prec, eloc, esz, mloc, msz, bias) == (0, 64, 52, 11, 0, 52, 1023), "IEEE double properties")
                    d.dtype = np.dtype("<f8")
                elif cls == 0:
                    need(not (b0 & 1), "little-endian integer")
                    d.dtype = np.dtype(f"<i{sz}" if b0 & 0x08 else f"<u{sz}")
                elif cls == 3:
                    d.dtype = np.dtype(f"S{sz}")
                else:
                    raise SpecError(f"unexpected datatype class {cls}")
            elif mtype == 0x05:
                ver, alloc, wtime, defined = self.u("BBBB", body)
                need(ver == 2 and alloc in (1, 2, 3) and wtime in (0, 1, 2), "fill value message v2 fields")
                if defined:
                    fsz = self.u("I", body + 4)[0]
                    d.fill = (fsz, self.buf[body + 8:body + 8 + fsz])
            elif mtype == 0x0B:
                ver, nf = self.u("BB", body)
                need(ver == 1 and nf >= 1, "filter pipeline version")
                p = body + 8
                for _ in range(nf):
                    fid, nlen, fl, ncd = self.u("HHHH", p)
                    p += 8
                    need(nlen % 8 == 0, "filter name length padded to 8")
                    p += nlen
                    cd = [self.u("I", p + 4 * i)[0] for i in range(ncd)]
                    p += 4 * ncd + (4 if ncd % 2 else 0)
                    d.filters.append((fid, cd))
                need(p <= body + size, "filter pipeline inside its message")
            elif mtype == 0x08:
                layout_body = body
        d.attrs = self.attributes(msgs)
        es = d.dtype.itemsize
        if d.fill is not None:
            need(d.fill[0] == es, "fill value size == element size")
        ver, cls = self.u("BB", layout_body)
        need(ver == 3, "layout version 3")
        rank = len(d.shape)
        if cls == 1:
            addr, size = self.u("QQ", layout_body + 2)
            need(not d.filters, "contiguous datasets cannot be filtered")
            need(size == int(np.prod(d.shape)) * es and addr + size <= self.n, "contiguous storage size / inside file")
            d.layout = "contiguous"
            d.data = np.frombuffer(self.buf, d.dtype, int(np.prod(d.shape)), addr).reshape(d.shape).copy()
            return d
        need(cls == 2, "layout class")
        nd1 = self.buf[layout_body + 2]
        need(nd1 == rank + 1, "chunked dimensionality must be rank + 1")
        bt = self.u("Q", layout_body + 3)[0]
        cd = [self.u("I", layout_body + 11 + 4 * i)[0] for i in range(nd1)]
        need(cd[-1] == es, "last chunk dimension == element size")
        need(all(c >= 1 for c in cd) and all(cd[i] <= max(d.shape[i], 1) for i in range(rank)), "chunk dims within the dataspace")
        d.layout, d.chunk = "chunked", tuple(cd[:-1])
        fillv = np.frombuffer(d.fill[1], d.dtype)[0] if d.fill is not None else np.zeros((), d.dtype)
        d.data = np.full(d.shape, fillv, d.dtype)
        if bt == UNDEF:
            return d
        keysz = 8 + 8 * nd1
        nodesz = 24 + 2 * CHUNK_K * (keysz + 8) + keysz
        raw_size = int(np.prod(d.chunk)) * es
        seen = []

        def key(p):
            size, mask = self.u("II", p)
            offs = tuple(self.u("Q", p + 8 + 8 * i)[0] for i in range(nd1))
            need(offs[-1] == 0, "element-size offset of a chunk key must be 0")
            return size, mask, offs[:-1]

        def node(addr, level_expected, left, right):
            need(addr % 8 == 0 and addr + nodesz <= self.n, "chunk B-tree node aligned and inside the file at full size")
            need(self.buf[addr:addr + 4] == b"TREE", "chunk B-tree node signature")
            ntype, level, nent, lsib, rsib = self.u("BBHQQ", addr + 4)
            need(ntype == 1 and 1 <= nent <= 2 * CHUNK_K, "chunk B-tree node type / entry count")
            need(level == level_expected, "chunk B-tree levels")
            need(lsib == left and rsib == right, "sibling addresses")
            keys = [key(addr + 24 + i * (keysz + 8)) for i in range(nent + 1)]
            kids = [self.u("Q", addr + 24 + i * (keysz + 8) + keysz)[0] for i in range(nent)]
            need(all(keys[i][2] < keys[i + 1][2] for i in range(nent)), "chunk keys must increase strictly (lexicographic offsets)")
            return level, keys, kids

        def descend(addr, level_expected, left, right):
            level, keys, kids = node(addr, level_expected, left, right)
            if level == 0:
                for (size, mask, offs), c in zip(keys, kids):
                    need(c + size <= self.n, "chunk inside file")
                    need(all(offs[i] % d.chunk[i] == 0 and offs[i] < d.shape[i] for i in range(rank)), "chunk aligned to the chunk grid, inside the dataspace")
                    raw = self.buf[c:c + size]
                    for i, (fid, cdv) in reversed(list(enumerate(d.filters))):
                        if mask & (1 << i):
                            continue
                        need(fid == 1, "only deflate expected")
                        raw = zlib.decompress(raw)
                    need(len(raw) == raw_size, f"chunk size after filters {len(raw)} != {raw_size}")
                    blk = np.frombuffer(raw, d.dtype).reshape(d.chunk)
                    sl_o = tuple(slice(offs[i], min(offs[i] + d.chunk[i], d.shape[i])) for i in range(rank))
                    sl_i = tuple(slice(0, s.stop - s.start) for s in sl_o)
                    d.data[sl_o] = blk[sl_i]
                    seen.append(offs)
                return keys[0], keys[-1]
            first = last = None
            # children of one node are siblings of each other; across nodes the chain continues (checked by
            # walking every level left to right below)
            for i, c in enumerate(kids):
                k0, k1 = descend_child(c, level - 1)
                need(k0[2] == keys[i][2], "internal key == first key of its child")
                first = first or k0
                last = k1
            need(keys[-1][2] == last[2], "right-most key of an internal node == right-most key of its last child")
            return keys[0], keys[-1]

        # sibling chains: collect the nodes of every level first
        levels = {}

        def collect(addr):
            ntype, level, nent = self.u("BBH", addr + 4)
            levels.setdefault(level, []).append(addr)
            if level > 0:
                for i in range(nent):
                    collect(self.u("Q", addr + 24 + i * (keysz + 8) + keysz)[0])

        collect(bt)
        sib = {}
        for level, nodes in levels.items():
            for i, a in enumerate(nodes):
                sib[a] = (nodes[i - 1] if i else UNDEF, nodes[i + 1] if i + 1 < len(nodes) else UNDEF)

        def descend_child(addr, level_expected):
            return descend(addr, level_expected, *sib[addr])

        root_level = self.buf[bt + 5]
        d.btree_levels = root_level + 1
        k0, k1 = descend(bt, root_level, UNDEF, UNDEF)
        need(seen == sorted(seen) and len(set(seen)) == len(seen), "chunks visited in key order, no duplicates")
        need(all(k1[2][i] >= seen[-1][i] for i in range(rank)) and k1[2] > seen[-1], "right-most key beyond the last chunk")
        d.nchunks = len(seen)
        return d


def walk(path):
    w = Walker(path)
    ohdr, btree, heap = w.superblock()
    out = {}
    entries, root_attrs = w.group(ohdr, btree, heap)
    out["/"] = root_attrs
    out["trailing_bytes"] = w.trailing
    for name, oh, scratch in entries:
        need(scratch is not None, "groups are linked with cached B-tree/heap addresses")
        sub, _ = w.group(oh, *scratch)
        for dname, doh, _ in sub:
            out[f"/{name}/{dname}"] = w.dataset(doh)
    return out
