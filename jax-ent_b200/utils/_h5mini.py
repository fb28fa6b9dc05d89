"""A small, dependency-free HDF5 writer / reader for optimisation histories.

The reference stores `OptimizationHistory` objects with h5py (`jaxent/src/utils/hdf.py:258-427`: nested groups, float / int
datasets, string / int / bool attributes) and its analysis scripts read them back with h5py.  h5py is not available in every
environment this package runs in (it is absent from the build image), so `utils/hdf.py` uses h5py when it imports and this
module otherwise.  It implements the subset of the HDF5 file format (version-0 superblock, version-1 object headers,
symbol-table groups: local heap + version-1 B-tree + symbol nodes) that `libver="earliest"` h5py files use:

  writing  groups, contiguous datasets of bool / int / uint / float (any shape, incl. scalars), attributes of the same
           types plus variable-length UTF-8 strings (what `attrs["x"] = "text"` produces under h5py, global heap) and
           numpy bools as the h5py (FALSE, TRUE) int8 enumeration;
  reading  the same, plus continuation blocks, compact and chunked datasets with the deflate / shuffle filters (the
           reference writes gzip-compressed chunked datasets), fixed-length strings, version 1-3 attribute messages.

The in-memory model mirrors the corner of the h5py API the (de)serialisers need: `Group.create_group`,
`Group.create_dataset`, `Group.attrs`, `group[path]`, `len(group)`, `dataset[()]`.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Optional

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16  # library defaults: 8 symbols per symbol node, 32 children per B-tree node


# ---------------------------------------------------------------------------------------------------------- model
class Dataset:
    def __init__(self, data, attrs: Optional[dict] = None):
        self.data = np.asarray(data)
        self.attrs: Dict[str, object] = dict(attrs or {})

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    def __getitem__(self, key):
        if key == () or key is Ellipsis:
            return self.data[()] if self.data.shape == () else self.data.copy()
        return self.data[key]

    def __array__(self, dtype=None, copy=None):
        return self.data.astype(dtype) if dtype is not None else self.data


class Group:
    def __init__(self):
        self.children: Dict[str, object] = {}
        self.attrs: Dict[str, object] = {}

    # -- h5py-like construction
    def create_group(self, path: str) -> "Group":
        node = self
        parts = [p for p in path.split("/") if p]
        for i, p in enumerate(parts):
            if p in node.children:
                if i == len(parts) - 1 or not isinstance(node.children[p], Group):
                    raise ValueError(f"unable to create group (name already exists): {path}")
                node = node.children[p]
            else:
                g = Group()
                node.children[p] = g
                node = g
        return node

    def create_dataset(self, path: str, data=None, **kwargs) -> Dataset:
        # compression / chunking options are accepted and ignored: the writer stores contiguous datasets
        parts = [p for p in path.split("/") if p]
        node = self
        for p in parts[:-1]:
            node = node.children[p] if p in node.children else node.create_group(p)
        if parts[-1] in node.children:
            raise ValueError(f"unable to create dataset (name already exists): {path}")
        ds = Dataset(np.asarray(data))
        node.children[parts[-1]] = ds
        return ds

    # -- h5py-like access
    def __getitem__(self, path: str):
        node = self
        for p in [q for q in path.split("/") if q]:
            if not isinstance(node, Group) or p not in node.children:
                raise KeyError(f"unable to open object (component not found): {path}")
            node = node.children[p]
        return node

    def __contains__(self, path: str) -> bool:
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __len__(self) -> int:
        return len(self.children)

    def __iter__(self):
        return iter(sorted(self.children))

    def keys(self):
        return sorted(self.children)

    def items(self):
        return [(k, self.children[k]) for k in sorted(self.children)]


class File(Group):
    """`with File(name, "w") as f:` writes on exit; `File(name, "r")` parses the whole file eagerly."""

    def __init__(self, filename: str, mode: str = "r"):
        super().__init__()
        if mode not in ("r", "w"):
            raise ValueError("mode must be 'r' or 'w'")
        self.filename, self.mode = filename, mode
        if mode == "r":
            with open(filename, "rb") as fh:
                root = _Reader(fh.read()).read()
            self.children, self.attrs = root.children, root.attrs

    def close(self):
        if self.mode == "w":
            with open(self.filename, "wb") as fh:
                fh.write(_Writer().write(self))
            self.mode = "closed"

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        return False


# ---------------------------------------------------------------------------------------------------------- datatypes
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    """datatype message body (version 1) of a numpy scalar type"""
    dt = np.dtype(dt)
    if dt.kind == "b":  # h5py's boolean: enumeration of int8 {FALSE = 0, TRUE = 1}
        base = _dtype_message(np.dtype("i1"))
        names = _pad8(b"FALSE\0") + _pad8(b"TRUE\0")
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        if dt.itemsize == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        elif dt.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        elif dt.itemsize == 2:
            sign, eloc, esize, msize, bias = 15, 10, 5, 10, 15
        else:
            raise TypeError(f"unsupported float size {dt.itemsize}")
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, sign, 0, dt.itemsize, 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, max(dt.itemsize, 1))
    raise TypeError(f"unsupported dtype {dt}")


_VLEN_STR_TYPE = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x10, 0, 0, 1)  # vlen of UTF-8 char


def _dataspace_message(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _to_array(value) -> np.ndarray:
    a = np.asarray(value)
    if a.dtype.kind == "O":
        raise TypeError("object arrays are not supported")
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return a


# ---------------------------------------------------------------------------------------------------------- writer
class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # the superblock is filled in last
        self.gheap_objects = []   # (bytes) of the single global heap collection (variable-length strings)
        self.gheap_patches = []   # (position of the 16-byte vlen descriptor in buf, object index)

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    # -- messages
    @staticmethod
    def _message(mtype: int, body: bytes) -> bytes:
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _attribute(self, name: str, value) -> bytes:
        nm = name.encode("utf-8") + b"\0"
        if isinstance(value, (str, bytes)) and not isinstance(value, np.ndarray):
            raw = value.encode("utf-8") if isinstance(value, str) else value
            self.gheap_objects.append(raw)
            dt, sp, data = _VLEN_STR_TYPE, _dataspace_message(()), ("vlen", len(raw), len(self.gheap_objects))
        else:
            a = _to_array(value)
            if a.dtype.kind == "U":
                raise TypeError("arrays of str are not supported as attributes")
            dt, sp, data = _dtype_message(a.dtype), _dataspace_message(a.shape), a.tobytes()
        head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp)
        return head, data

    def _object_header(self, messages) -> int:
        """messages: list of (type, body) or (type, head, vlen-descriptor) for attributes holding a string"""
        out, patches = b"", []
        for m in messages:
            if len(m) == 2:
                out += self._message(m[0], m[1])
            else:
                mtype, head, (_, length, index) = m
                body = _pad8(head + b"\0" * 16)
                patches.append((len(out) + 8 + len(head), length, index))
                out += struct.pack("<HHB3x", mtype, len(body), 0) + body
        addr = self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(out)) + out)
        for off, length, index in patches:
            self.gheap_patches.append((addr + 16 + off, length, index))
        return addr

    def _attr_messages(self, attrs: dict):
        msgs = []
        for k in attrs:
            head, data = self._attribute(k, attrs[k])
            msgs.append((0x000C, head, data) if isinstance(data, tuple) else (0x000C, head + data))
        return msgs

    # -- objects
    def _dataset(self, ds: Dataset) -> int:
        a = _to_array(ds.data)
        raw = a.tobytes()  # C order, also for 0-d and non-contiguous inputs
        data_addr = self.alloc(raw) if raw else UNDEF
        msgs = [(0x0001, _dataspace_message(a.shape)), (0x0003, _dtype_message(a.dtype)), (0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),
                (0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw)))]
        return self._object_header(msgs + self._attr_messages(ds.attrs))

    def _group(self, g: Group):
        """returns (object header address, B-tree address, heap address)"""
        names = sorted(g.children, key=lambda s: s.encode("utf-8"))
        child_addr = {}
        for n in names:
            c = g.children[n]
            child_addr[n] = self._group(c)[0] if isinstance(c, Group) else self._dataset(c)
        # local heap: offset 0 is the empty string, names follow, a free block closes the segment
        seg, name_off = bytearray(8), {}
        for n in names:
            name_off[n] = len(seg)
            seg += _pad8(n.encode("utf-8") + b"\0")
        free_off = len(seg)
        free_size = max(16, 88 - len(seg) if len(seg) < 72 else 16)
        seg += struct.pack("<QQ", 1, free_size) + b"\0" * (free_size - 16)  # next = 1: end of the free list
        seg_addr = self.alloc(bytes(seg))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        # symbol nodes
        per = 2 * LEAF_K
        leaves = []  # (address, largest name)
        for i in range(0, len(names), per):
            chunk = names[i:i + per]
            ent = b"".join(struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0) for n in chunk)
            ent += b"\0" * (40 * (per - len(chunk)))
            leaves.append((self.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)) + ent), chunk[-1]))
        # B-tree levels
        level, nodes = 0, leaves
        fan = 2 * INTERNAL_K
        while True:
            groups = [nodes[i:i + fan] for i in range(0, len(nodes), fan)] or [[]]
            addrs = [self.alloc(b"\0" * (24 + (2 * fan + 1) * 8)) for _ in groups]
            new_nodes = []
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                first_key = name_off[groups[gi - 1][-1][1]] if gi > 0 else 0
                body = b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right) + struct.pack("<Q", first_key)
                for addr, last in grp:
                    body += struct.pack("<QQ", addr, name_off[last])
                self.buf[addrs[gi]:addrs[gi] + len(body)] = body
                new_nodes.append((addrs[gi], grp[-1][1] if grp else None))
            if len(new_nodes) == 1:
                btree_addr = new_nodes[0][0]
                break
            nodes, level = new_nodes, level + 1
        hdr = self._object_header([(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + self._attr_messages(g.attrs))
        return hdr, btree_addr, heap_addr

    def write(self, root: Group) -> bytes:
        hdr, btree, heap = self._group(root)
        if self.gheap_objects:
            body = b""
            for i, raw in enumerate(self.gheap_objects):
                body += struct.pack("<HH4xQ", i + 1, 1, len(raw)) + _pad8(raw)
            size = max(4096, 16 + len(body) + 16)
            size += -size % 8
            free = size - 16 - len(body)
            col = b"GCOL" + struct.pack("<B3xQ", 1, size) + body + struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)
            col_addr = self.alloc(col)
            for pos, length, index in self.gheap_patches:
                self.buf[pos:pos + 16] = struct.pack("<IQI", length, col_addr, index)
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


# ---------------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.base = 0

    def u(self, fmt: str, off: int):
        return struct.unpack_from("<" + fmt, self.d, off)

    def read(self) -> Group:
        off = 0
        while self.d[off:off + 8] != SIGNATURE:
            off = 512 if off == 0 else off * 2
            if off >= len(self.d):
                raise ValueError("not an HDF5 file")
        ver = self.d[off + 8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver}: only the 'earliest' file format (versions 0 and 1) is supported")
        so, sl = self.d[off + 13], self.d[off + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = self.u("Q", p)[0]
        root_entry = p + 32
        hdr_addr = self.u("Q", root_entry + 8)[0]
        return self._object(hdr_addr)

    # -- object headers
    def _messages(self, addr: int):
        a = self.base + addr
        ver, _, nmsg, _, size = self.u("BBHII", a)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks, out = [(a + 16, size)], []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = self.u("HHB", p)
                body = self.d[p + 8:p + 8 + msize]
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from("<QQ", body)
                    blocks.append((self.base + coff, clen))
                out.append((mtype, body))
                p += 8 + msize
        return out

    def _object(self, addr: int):
        msgs = self._messages(addr)
        types = {m[0] for m in msgs}
        if 0x0011 in types:
            node = Group()
            btree, heap = struct.unpack_from("<QQ", next(b for t, b in msgs if t == 0x0011))
            for name, child in self._symbols(btree, heap):
                node.children[name] = self._object(child)
        elif 0x0008 in types:
            node = Dataset(self._dataset(msgs))
        else:
            raise NotImplementedError("object is neither a symbol-table group nor a dataset (new-style groups are not supported)")
        for t, b in msgs:
            if t == 0x000C:
                k, v = self._attribute(b)
                node.attrs[k] = v
        return node

    # -- groups
    def _heap_string(self, heap_addr: int, off: int) -> str:
        a = self.base + heap_addr
        assert self.d[a:a + 4] == b"HEAP", "bad local heap signature"
        seg = self.base + self.u("Q", a + 24)[0]
        end = self.d.index(b"\0", seg + off)
        return self.d[seg + off:end].decode("utf-8")

    def _symbols(self, btree: int, heap: int):
        a = self.base + btree
        sig = self.d[a:a + 4]
        if sig == b"SNOD":
            n = self.u("H", a + 6)[0]
            for i in range(n):
                name_off, obj = self.u("QQ", a + 8 + 40 * i)
                yield self._heap_string(heap, name_off), obj
            return
        assert sig == b"TREE", "bad B-tree signature"
        _ntype, _level, used = self.u("BBH", a + 4)
        for i in range(used):
            child = self.u("Q", a + 24 + 8 + 16 * i)[0]
            yield from self._symbols(child, heap)

    # -- datatypes
    def _dtype(self, b: bytes, off: int = 0):
        """returns (numpy dtype or ('vlen_str',) / ('enum', base, names, values), message length)"""
        cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", b, off)
        cls, ver = cv & 0x0F, cv >> 4
        p = off + 8
        if cls == 0:
            order = ">" if b0 & 1 else "<"
            return np.dtype(f"{order}{'i' if b0 & 8 else 'u'}{size}"), p + 4 - off
        if cls == 1:
            order = ">" if b0 & 1 else "<"
            return np.dtype(f"{order}f{size}"), p + 12 - off
        if cls == 3:
            return np.dtype(f"S{size}"), p - off
        if cls == 8:
            nmemb = b0 | (b1 << 8)
            base, blen = self._dtype(b, p)
            p += blen
            names = []
            for _ in range(nmemb):
                end = b.index(b"\0", p)
                names.append(b[p:end].decode())
                p = p + (((end - p) // 8 + 1) * 8 if ver < 3 else end - p + 1)
            values = np.frombuffer(b, dtype=base, count=nmemb, offset=p)
            return ("enum", base, names, values.copy()), p + nmemb * base.itemsize - off
        if cls == 9:
            if (b0 & 0x0F) != 1:
                raise NotImplementedError("variable-length sequences")
            _, blen = self._dtype(b, p)
            return ("vlen_str",), p + blen - off
        raise NotImplementedError(f"datatype class {cls}")

    def _dataspace(self, b: bytes):
        ver, rank, flags = struct.unpack_from("<BBB", b)
        p = 8 if ver == 1 else 4
        if ver == 2 and b[3] == 2:
            return None  # null dataspace
        return tuple(struct.unpack_from("<Q", b, p + 8 * i)[0] for i in range(rank))

    def _global_heap_object(self, col_addr: int, index: int) -> bytes:
        a = self.base + col_addr
        assert self.d[a:a + 4] == b"GCOL", "bad global heap signature"
        size = self.u("Q", a + 8)[0]
        p, end = a + 16, a + size
        while p + 16 <= end:
            idx, _ref, osize = self.u("HH4xQ", p)
            if idx == 0:
                break
            if idx == index:
                return self.d[p + 16:p + 16 + osize]
            p += 16 + osize + (-osize % 8)
        raise KeyError(f"global heap object {index} not found")

    def _decode(self, dt, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, tuple) and dt[0] == "vlen_str":
            out = []
            for i in range(n):
                length, col, idx = struct.unpack_from("<IQI", raw, 16 * i)
                out.append(self._global_heap_object(col, idx)[:length].decode("utf-8") if length else "")
            return out[0] if shape == () else np.array(out, dtype=object).reshape(shape)
        if isinstance(dt, tuple) and dt[0] == "enum":
            _, base, names, values = dt
            a = np.frombuffer(raw, dtype=base, count=n).reshape(shape)
            if sorted(names) == ["FALSE", "TRUE"]:
                a = a == values[names.index("TRUE")]
            return a[()] if shape == () else a.copy()
        a = np.frombuffer(raw, dtype=dt, count=n).reshape(shape)
        if dt.byteorder == ">":
            a = a.astype(dt.newbyteorder("<"))
        if shape == ():
            v = a[()]
            return v.decode("utf-8", "replace") if dt.kind == "S" else v
        return a.copy()

    def _attribute(self, b: bytes):
        ver = b[0]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            p = 8
            name = b[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += nsz + (-nsz % 8)
            dt, _ = self._dtype(b, p)
            p += dsz + (-dsz % 8)
            shape = self._dataspace(b[p:p + ssz])
            p += ssz + (-ssz % 8)
        elif ver in (2, 3):
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = b[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += nsz
            dt, _ = self._dtype(b, p)
            p += dsz
            shape = self._dataspace(b[p:p + ssz])
            p += ssz
        else:
            raise NotImplementedError(f"attribute message version {ver}")
        if shape is None:
            return name, None
        return name, self._decode(dt, shape, b[p:])

    # -- datasets
    def _dataset(self, msgs):
        body = {t: b for t, b in msgs}
        dt, _ = self._dtype(body[0x0003])
        shape = self._dataspace(body[0x0001])
        lay = body[0x0008]
        if lay[0] not in (1, 2, 3):
            raise NotImplementedError(f"data layout message version {lay[0]}")
        itemsize = 16 if isinstance(dt, tuple) and dt[0] == "vlen_str" else (dt[1].itemsize if isinstance(dt, tuple) else dt.itemsize)
        nbytes = (int(np.prod(shape)) if shape else 1) * itemsize
        if lay[0] < 3:  # versions 1 and 2 (files written by old libraries): dimensionality, class, 5 reserved bytes, address, sizes
            ndim, cls = lay[1], lay[2]
            if cls == 1:
                addr = struct.unpack_from("<Q", lay, 8)[0]
                raw = b"\0" * nbytes if addr == UNDEF else self.d[self.base + addr:self.base + addr + nbytes]
            elif cls == 0:
                size = struct.unpack_from("<I", lay, 8 + 4 * ndim)[0]
                raw = lay[12 + 4 * ndim:12 + 4 * ndim + size]
            else:
                raise NotImplementedError("chunked datasets with a version-1/2 layout message")
            return self._decode(dt, shape, raw)
        cls = lay[1]
        if cls == 0:
            size = struct.unpack_from("<H", lay, 2)[0]
            raw = lay[4:4 + size]
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", lay, 2)
            raw = b"\0" * nbytes if addr == UNDEF else self.d[self.base + addr:self.base + addr + size]
        elif cls == 2:
            raw = self._chunked(lay, body.get(0x000B), shape, itemsize)
        else:
            raise NotImplementedError(f"data layout class {cls}")
        return self._decode(dt, shape, raw)

    def _filters(self, b: Optional[bytes]):
        if b is None:
            return []
        ver, n = b[0], b[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", b, p)[0]
            if ver == 1 or fid >= 256:
                namelen, _flags, ncd = struct.unpack_from("<HHH", b, p + 2)
                p += 8 + namelen + (-namelen % 8 if ver == 1 else 0)
            else:
                _flags, ncd = struct.unpack_from("<HH", b, p + 2)
                p += 6
            cd = struct.unpack_from(f"<{ncd}I", b, p)
            p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def _chunked(self, lay: bytes, pipeline: Optional[bytes], shape, itemsize: int) -> bytes:
        ndim = lay[2]  # rank + 1
        btree = struct.unpack_from("<Q", lay, 3)[0]
        cdims = struct.unpack_from(f"<{ndim}I", lay, 11)[:-1]
        filters = self._filters(pipeline)
        out = np.zeros(shape, dtype=f"V{itemsize}")
        if btree == UNDEF:
            return out.tobytes()

        def walk(addr):
            a = self.base + addr
            assert self.d[a:a + 4] == b"TREE", "bad chunk B-tree signature"
            _ntype, level, used = self.u("BBH", a + 4)
            ksize = 8 + 8 * ndim
            p = a + 24
            for _ in range(used):
                csize, mask = self.u("II", p)
                offs = self.u(f"{ndim}Q", p + 8)[:-1]
                child = self.u("Q", p + ksize)[0]
                if level > 0:
                    walk(child)
                else:
                    raw = self.d[self.base + child:self.base + child + csize]
                    for i, (fid, cd) in reversed(list(enumerate(filters))):
                        if mask & (1 << i):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            n = len(raw) // itemsize
                            raw = np.frombuffer(raw, np.uint8).reshape(itemsize, n).T.tobytes()
                        else:
                            raise NotImplementedError(f"filter {fid}")
                    chunk = np.frombuffer(raw, dtype=f"V{itemsize}", count=int(np.prod(cdims))).reshape(cdims)
                    sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
                p += ksize + 8

        walk(btree)
        return out.tobytes()
