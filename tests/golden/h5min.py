"""TEST INFRASTRUCTURE ONLY (used by tests/golden/make_mask_golden.py, nothing else imports it).

Minimal read-only HDF5 parser - just enough for the two netCDF4 example files the reference ships
(examples/ngwerere/ngwerere_piv.nc, ngwerere_masked.nc: superblock v2, version-2 object headers with continuation blocks,
contiguous and chunked datasets indexed by a v1 B-tree, shuffle + deflate filters, compact attribute messages; dense link /
attribute storage is searched by pattern instead of walking the fractal heaps).  There is no h5py / netCDF4 / xarray in this
image; the HDF5 file format specification (version 3.0) is public."""
import struct, zlib
import numpy as np

class H5:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n" and self.b[8] in (2, 3)
        self.objs = {}
        pos = 0
        while True:
            pos = self.b.find(b"OHDR", pos)
            if pos < 0: break
            try:
                self.objs[pos] = self.parse_ohdr(pos)
            except Exception as e:
                self.objs[pos] = {"error": repr(e)}
            pos += 4

    def parse_ohdr(self, pos):
        b = self.b
        assert b[pos:pos+4] == b"OHDR" and b[pos+4] == 2
        flags = b[pos+5]
        p = pos + 6
        if flags & 0x20: p += 16
        if flags & 0x10: p += 4
        nsz = 1 << (flags & 3)
        size0 = int.from_bytes(b[p:p+nsz], "little"); p += nsz
        obj = {"msgs": [], "flags": flags}
        self.parse_msgs(obj, p, p + size0, flags)
        return obj

    def parse_msgs(self, obj, p, end, flags):
        b = self.b
        while p + 4 <= end:
            mtype = b[p]; msize = struct.unpack("<H", b[p+1:p+3])[0]; mflags = b[p+3]; p += 4
            if flags & 0x04: p += 2
            data = b[p:p+msize]
            if mtype == 0x10:
                off, ln = struct.unpack("<QQ", data[:16])
                assert b[off:off+4] == b"OCHK"
                self.parse_msgs(obj, off + 4, off + ln - 4, flags)
            elif mtype != 0:
                obj["msgs"].append((mtype, mflags, data))
            p += msize

    # ---- message decoders ------------------------------------------------------------------------------------------
    @staticmethod
    def dataspace(d):
        ver, rank, fl = d[0], d[1], d[2]
        p = 4 if ver == 2 else 8
        dims = struct.unpack("<%dQ" % rank, d[p:p+8*rank])
        return tuple(dims)

    @staticmethod
    def datatype(d):
        cls = d[0] & 0x0F; bits = d[1:4]; size = struct.unpack("<I", d[4:8])[0]
        if cls == 0:
            signed = bool(bits[0] & 0x08)
            return np.dtype(("<i" if signed else "<u") + str(size)), 8 + 4
        if cls == 1:
            return np.dtype("<f" + str(size)), 8 + 12
        if cls == 3:
            return np.dtype("S" + str(size)), 8
        if cls == 9:   # variable length
            return ("vlen", d[1] & 0x0F), None
        if cls == 7:   # reference
            return np.dtype("V" + str(size)), 8
        if cls == 6:   # compound
            return np.dtype("V" + str(size)), None
        if cls == 8:   # enum
            return np.dtype("V" + str(size)), None
        return np.dtype("V" + str(size)), None

    def attribute(self, d):
        ver = d[0]
        assert ver in (1, 2, 3), ver
        if ver == 1:
            nsz, tsz, ssz = struct.unpack("<HHH", d[2:8]); p = 8
            pad = lambda n: (n + 7) & ~7
            name = d[p:p+nsz].split(b"\0")[0].decode(); p += pad(nsz)
            dt = d[p:p+tsz]; p += pad(tsz); ds = d[p:p+ssz]; p += pad(ssz)
        else:
            nsz, tsz, ssz = struct.unpack("<HHH", d[2:8]); p = 8
            if ver == 3: p += 1
            name = d[p:p+nsz].split(b"\0")[0].decode(); p += nsz
            dt = d[p:p+tsz]; p += tsz; ds = d[p:p+ssz]; p += ssz
        dtype, _ = self.datatype(dt)
        shape = self.dataspace(ds) if len(ds) >= 4 and ds[1] > 0 else ()
        raw = d[p:]
        if isinstance(dtype, tuple):   # vlen string: (length 4, global heap address 8, index 4)
            try:
                ln, addr, idx = struct.unpack("<IQI", raw[:16])
                return name, self.gheap(addr, idx)[:ln]
            except Exception as e:
                return name, ("vlen?", repr(e))
        n = int(np.prod(shape)) if shape else 1
        if dtype.kind == "V":
            return name, raw[:dtype.itemsize * n]
        val = np.frombuffer(raw[:dtype.itemsize * n], dtype=dtype)
        if dtype.kind == "S":
            return name, val[0].split(b"\0")[0].decode(errors="replace") if n == 1 else [v.decode(errors="replace") for v in val]
        return name, (val[0] if not shape else val.reshape(shape))

    def gheap(self, addr, idx):
        b = self.b
        assert b[addr:addr+4] == b"GCOL"
        size = struct.unpack("<Q", b[addr+8:addr+16])[0]
        p = addr + 16
        while p < addr + size:
            i, ref, _, osz = struct.unpack("<HHIQ", b[p:p+16])
            if i == idx: return b[p+16:p+16+osz]
            if i == 0: break
            p += 16 + ((osz + 7) & ~7)
        raise KeyError(idx)

    def describe(self, pos):
        o = self.objs[pos]
        out = {"attrs": {}}
        for mtype, mflags, d in o.get("msgs", []):
            if mtype == 0x01: out["shape"] = self.dataspace(d)
            elif mtype == 0x03: out["dtype"] = self.datatype(d)[0]
            elif mtype == 0x08: out["layout"] = d
            elif mtype == 0x0B: out["filters"] = self.filters(d)
            elif mtype == 0x0C:
                try:
                    k, v = self.attribute(d); out["attrs"][k] = v
                except Exception as e:
                    out["attrs"]["?%d" % len(out["attrs"])] = repr(e)
            elif mtype == 0x06: out.setdefault("links", []).append(self.link(d))
            elif mtype == 0x15: out["attr_info"] = d
            elif mtype == 0x02: out["link_info"] = d
        return out

    @staticmethod
    def link(d):
        ver, fl = d[0], d[1]; p = 2
        ltype = 0
        if fl & 0x08: ltype = d[p]; p += 1
        if fl & 0x04: p += 8
        if fl & 0x10: p += 1
        lsz = 1 << (fl & 3)
        ln = int.from_bytes(d[p:p+lsz], "little"); p += lsz
        name = d[p:p+ln].decode(); p += ln
        addr = struct.unpack("<Q", d[p:p+8])[0] if ltype == 0 else None
        return name, addr

    @staticmethod
    def filters(d):
        ver, n = d[0], d[1]
        p = 2 if ver == 2 else 8
        out = []
        for _ in range(n):
            fid = struct.unpack("<H", d[p:p+2])[0]; p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = struct.unpack("<H", d[p:p+2])[0]; p += 2
            fl, ncv = struct.unpack("<HH", d[p:p+4]); p += 4
            if nlen: p += (nlen + 7) & ~7 if ver == 1 else nlen
            cv = struct.unpack("<%dI" % ncv, d[p:p+4*ncv]); p += 4 * ncv
            if ver == 1 and ncv % 2: p += 4
            out.append((fid, cv))
        return out

    def read(self, pos):
        o = self.describe(pos)
        shape, dtype, lay = o["shape"], o["dtype"], o["layout"]
        ver, cls = lay[0], lay[1]
        assert ver == 3, ("layout version", ver)
        n = int(np.prod(shape)) if shape else 1
        if cls == 0:      # compact
            sz = struct.unpack("<H", lay[2:4])[0]
            return np.frombuffer(lay[4:4+sz], dtype=dtype).reshape(shape)
        if cls == 1:      # contiguous
            addr, sz = struct.unpack("<QQ", lay[2:18])
            if addr == 0xFFFFFFFFFFFFFFFF: return np.zeros(shape, dtype)
            return np.frombuffer(self.b[addr:addr+dtype.itemsize*n], dtype=dtype).reshape(shape)
        assert cls == 2
        rank1 = lay[2]; bt = struct.unpack("<Q", lay[3:11])[0]
        cdims = struct.unpack("<%dI" % rank1, lay[11:11+4*rank1])
        chunk = cdims[:-1]
        out = np.zeros(shape, dtype)
        filt = o.get("filters", [])
        for key_off, size, mask, addr in self.btree_chunks(bt, rank1):
            raw = self.b[addr:addr+size]
            for j, (fid, cv) in reversed(list(enumerate(filt))):
                if mask & (1 << j): continue
                if fid == 1: raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cv[0]; a = np.frombuffer(raw, np.uint8); m = len(a) // es
                    raw = a[:m*es].reshape(es, m).T.tobytes() + a[m*es:].tobytes()
                elif fid == 3:   # fletcher32: strip checksum
                    raw = raw[:-4]
                else: raise NotImplementedError(("filter", fid))
            c = np.frombuffer(raw, dtype=dtype)[: int(np.prod(chunk))].reshape(chunk)
            sl = tuple(slice(k, min(k + cs, s)) for k, cs, s in zip(key_off, chunk, shape))
            out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def btree_chunks(self, addr, rank1):
        b = self.b
        assert b[addr:addr+4] == b"TREE", addr
        ntype, level, used = b[addr+4], b[addr+5], struct.unpack("<H", b[addr+6:addr+8])[0]
        assert ntype == 1
        p = addr + 24
        keysz = 8 + 8 * rank1
        for i in range(used):
            size, mask = struct.unpack("<II", b[p:p+8])
            offs = struct.unpack("<%dQ" % rank1, b[p+8:p+8+8*rank1])
            child = struct.unpack("<Q", b[p+keysz:p+keysz+8])[0]
            if level == 0: yield offs[:-1], size, mask, child
            else: yield from self.btree_chunks(child, rank1)
            p += keysz + 8


def dataset_names(h):
    """name -> object-header address of every dataset: the root group's links live in a fractal heap (dense storage); a hard-link
    message ends with the 8-byte address of the object header, directly after the link name, so the names are found by
    searching for the known header addresses."""
    import re

    names = {}
    for a in h.objs:
        for m in re.finditer(re.escape(struct.pack("<Q", a)), h.b):
            txt = re.findall(rb"[A-Za-z_0-9]+$", h.b[max(0, m.start() - 24):m.start()])
            if txt:
                names[txt[0].decode()] = a
    return names


def dense_attributes(h, key):
    """Every version-3 attribute message named `key` anywhere in the file (the variables' attributes are in dense storage):
    list of values in file order."""
    import re

    out = []
    k = key.encode() + b"\0"
    for m in re.finditer(re.escape(k), h.b):
        p = m.start() - 9
        if p >= 0 and h.b[p] == 3:
            try:
                out.append(h.attribute(h.b[p:p + 256])[1])
            except Exception:
                pass
    return out
