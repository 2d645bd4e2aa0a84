"""Minimal pure-python HDF5 reader (no libhdf5 / h5py in this image).

Reads the subset of HDF5 the Magnetizer input file uses (reference
python/prepare_input.py:93-150 writes it with h5py defaults; the Fortran side reads it
in source/IO_hdf5.f90:365-460): superblock v0, v1 object headers, symbol-table groups,
contiguous or chunked (v1 B-tree) datasets, optional deflate/shuffle filters, scalar
attributes.  Enough for `Input/*` and the root attributes
'Number of galaxies' / 'Number of snapshots' (IO_hdf5.f90:103-105).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Dataset:
    def __init__(self, f, shape, dtype, layout, filters, attrs, fill=None):
        self._f, self.shape, self.dtype, self._layout, self._filters, self.attrs = f, shape, dtype, layout, filters, attrs
        self._fill = fill

    def _empty(self):
        out = np.zeros(self.shape, self.dtype)
        if self._fill is not None and len(self._fill) == np.dtype(self.dtype).itemsize:
            out[...] = np.frombuffer(self._fill, self.dtype)[0]
        return out

    def read(self) -> np.ndarray:
        f = self._f
        kind = self._layout[0]
        if kind == "contiguous":
            _, addr, size = self._layout
            n = int(np.prod(self.shape)) if self.shape else 1
            if addr == _UNDEF:
                return self._empty()
            return np.frombuffer(f.buf, self.dtype, n, addr).reshape(self.shape).copy()
        if kind == "compact":
            return np.frombuffer(self._layout[1], self.dtype).reshape(self.shape).copy()
        _, btree, cdims = self._layout
        out = self._empty()
        if btree == _UNDEF:
            return out
        nd = len(self.shape)
        for offs, addr, size, mask in f._chunk_iter(btree, nd):
            raw = f.buf[addr:addr + size]
            for i, (fid, cd) in reversed(list(enumerate(self._filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else self.dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    raw = a.reshape(es, -1).T.tobytes()
                elif fid == 3:  # fletcher32: strip checksum
                    raw = raw[:-4]
                else:
                    raise NotImplementedError(f"HDF5 filter {fid}")
            chunk = np.frombuffer(raw, self.dtype, int(np.prod(cdims))).reshape(cdims)
            sl_out, sl_in = [], []
            for d in range(nd):
                lo = offs[d]
                hi = min(lo + cdims[d], self.shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
        return out


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        ver = self.buf[8]
        if ver != 0:
            raise NotImplementedError(f"superblock version {ver}")
        so, sl = self.buf[13], self.buf[14]
        if (so, sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths")
        # sig 8 + versions/sizes 8 + K values 4 + flags 4 + 4 addresses 32 = 56
        self.root = self._read_symbol_entry(56)

    # ------------------------------------------------------------------ low level
    def _u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.buf, off)

    def _read_symbol_entry(self, off):
        name_off, ohdr, cache, _ = self._u("QQII", off)
        scratch = self.buf[off + 24:off + 40]
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, scratch=scratch)

    def _messages(self, addr):
        ver, _, nmsg, _ref, hsize = self._u("BBHII", addr)
        if ver != 1:
            raise NotImplementedError(f"object header v{ver}")
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(msgs) < nmsg:
                mtype, msize, mflags = self._u("HHB", off)
                body = off + 8
                if mtype == 0x10:
                    coff, clen = self._u("QQ", body)
                    blocks.append((coff, clen))
                msgs.append((mtype, body, msize, mflags))
                off = body + msize
        return msgs

    def _parse_dtype(self, off):
        cv = self.buf[off]
        cls, bits0 = cv & 0x0F, self.buf[off + 1]
        size = self._u("I", off + 4)[0]
        if cls == 1:
            return np.dtype("<f%d" % size) if not (bits0 & 1) else np.dtype(">f%d" % size)
        if cls == 0:
            signed = bool(self.buf[off + 1] & 0x08)
            return np.dtype(("<i%d" if signed else "<u%d") % size)
        if cls == 3:
            return np.dtype("S%d" % size)
        raise NotImplementedError(f"datatype class {cls}")

    def _parse_dataspace(self, off):
        ver, rank, flags = self.buf[off], self.buf[off + 1], self.buf[off + 2]
        if ver == 1:
            p = off + 8
        elif ver == 2:
            p = off + 4
        else:
            raise NotImplementedError(f"dataspace v{ver}")
        return tuple(self._u("Q" * rank, p)) if rank else ()

    def _parse_attr(self, off):
        ver, _, nsz, dsz, ssz = self._u("BBHHH", off)
        if ver != 1:
            raise NotImplementedError(f"attribute v{ver}")
        pad = lambda n: (n + 7) & ~7
        p = off + 8
        name = self.buf[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        dt = self._parse_dtype(p)
        p += pad(dsz)
        shape = self._parse_dataspace(p)
        p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(self.buf, dt, n, p)
        if dt.kind == "S":
            return name, val[0].split(b"\0")[0].decode(errors="replace")
        return name, (val.reshape(shape).copy() if shape else val[0])

    def _chunk_iter(self, addr, nd):
        sig = self.buf[addr:addr + 4]
        if sig != b"TREE":
            raise ValueError("bad chunk B-tree")
        ntype, level, nent = self._u("BBH", addr + 4)
        p = addr + 24
        keysz = 8 + 8 * (nd + 1)
        for i in range(nent):
            csize, mask = self._u("II", p)
            offs = self._u("Q" * (nd + 1), p + 8)
            child = self._u("Q", p + keysz)[0]
            if level == 0:
                yield offs[:nd], child, csize, mask
            else:
                yield from self._chunk_iter(child, nd)
            p += keysz + 8

    def _group_entries(self, btree, heap):
        hdata = self._u("Q", heap + 24)[0]  # HEAP: sig4 ver1 res3 size8 free8 addr8

        def walk(addr):
            sig = self.buf[addr:addr + 4]
            if sig == b"TREE":
                _, level, nent = self._u("BBH", addr + 4)
                p = addr + 24 + 8  # skip first key
                for _ in range(nent):
                    child = self._u("Q", p)[0]
                    yield from walk(child)
                    p += 16
            elif sig == b"SNOD":
                nsym = self._u("H", addr + 6)[0]
                for i in range(nsym):
                    e = self._read_symbol_entry(addr + 8 + 40 * i)
                    s = hdata + e["name_off"]
                    name = self.buf[s:self.buf.index(b"\0", s)].decode()
                    yield name, e
            else:
                raise ValueError("bad group node")

        return dict(walk(btree))

    # ------------------------------------------------------------------ public
    def _open(self, ohdr):
        msgs = self._messages(ohdr)
        attrs, shape, dtype, layout, filters, stab, fill = {}, None, None, None, [], None, None
        for mtype, body, msize, _ in msgs:
            if mtype == 0x05:   # fill value (versions 1-3): what never-written chunks read as
                ver = self.buf[body]
                if ver in (1, 2) and (ver == 1 or self.buf[body + 3]):
                    sz = self._u("I", body + 4)[0]
                    fill = bytes(self.buf[body + 8:body + 8 + sz])
                elif ver == 3 and self.buf[body + 1] & 0x20:
                    sz = self._u("I", body + 2)[0]
                    fill = bytes(self.buf[body + 6:body + 6 + sz])
            elif mtype == 0x01:
                shape = self._parse_dataspace(body)
            elif mtype == 0x03:
                dtype = self._parse_dtype(body)
            elif mtype == 0x08:
                ver = self.buf[body]
                if ver != 3:
                    raise NotImplementedError(f"layout v{ver}")
                cls = self.buf[body + 1]
                if cls == 1:
                    a, s = self._u("QQ", body + 2)
                    layout = ("contiguous", a, s)
                elif cls == 2:
                    nd1 = self.buf[body + 2]
                    bt = self._u("Q", body + 3)[0]
                    dims = self._u("I" * nd1, body + 11)
                    layout = ("chunked", bt, tuple(dims[:-1]))
                else:
                    sz = self._u("H", body + 2)[0]
                    layout = ("compact", self.buf[body + 4:body + 4 + sz])
            elif mtype == 0x0B:
                ver, nf = self.buf[body], self.buf[body + 1]
                p = body + 8 if ver == 1 else body + 2
                for _ in range(nf):
                    fid, nlen, _fl, ncd = self._u("HHHH", p)
                    p += 8
                    if ver == 1 or fid >= 256:
                        p += (nlen + 7) & ~7 if ver == 1 else nlen
                    cd = self._u("I" * ncd, p)
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
            elif mtype == 0x0C:
                try:
                    k, v = self._parse_attr(body)
                    attrs[k] = v
                except NotImplementedError:
                    pass  # e.g. variable-length string attributes (Units/Description); not needed
            elif mtype == 0x11:
                stab = self._u("QQ", body)
        if stab is not None:
            return ("group", self._group_entries(*stab), attrs)
        return ("dataset", H5Dataset(self, shape, dtype, layout, filters, attrs, fill), attrs)

    def attrs(self, path="/"):
        return self._resolve(path)[2]

    def _resolve(self, path):
        node = self._open(self.root["ohdr"])
        for part in [p for p in path.split("/") if p]:
            if node[0] != "group":
                raise KeyError(path)
            node = self._open(node[1][part]["ohdr"])
        return node

    def keys(self, path="/"):
        node = self._resolve(path)
        return sorted(node[1].keys()) if node[0] == "group" else []

    def __getitem__(self, path) -> np.ndarray:
        node = self._resolve(path)
        if node[0] != "dataset":
            raise KeyError(f"{path} is a group")
        return node[1].read()


#: the 13 per-galaxy time series in the order of load_input_parameters
#: (reference source/input_parameters.f90:158-170)
INPUT_COLUMNS = ("r_disk", "v_disk", "r_bulge", "v_bulge", "r_halo", "v_halo", "nfw_cs1", "Mgas_disk",
                 "Mstars_disk", "SFR", "Mhalo", "Mstars_bulge", "central")


def load_magnetizer_input(path):
    """Returns (t_gyr[nz], inputs[13, ngal, nz]) from a Magnetizer input file."""
    f = H5File(path)
    t = np.asarray(f["Input/t"], np.float64).reshape(-1)
    cols = [np.asarray(f["Input/" + c], np.float64) for c in INPUT_COLUMNS]
    inputs = np.stack(cols, 0)
    at = f.attrs("/")
    ngal, nz = int(at.get("Number of galaxies", inputs.shape[1])), int(at.get("Number of snapshots", inputs.shape[2]))
    assert inputs.shape == (13, ngal, nz), (inputs.shape, ngal, nz)
    return t, np.ascontiguousarray(inputs)
