"""h5lite -- a minimal pure-NumPy HDF5 reader and writer for Keras 2.x weight files (models/<name><n>.h5).

The reference loads and saves its value network with Keras (`load_model(name + '.h5')`, `v_net.save('models/' + name + '.h5')`,
code/utils/alpha_nnet.py:11-12, 108-109).  h5py / libhdf5 are not available in this image, so this module implements the
small part of the HDF5 file format that those files use (HDF5 File Format Specification version 1.1/2.0):

  reader   superblock versions 0-1, version-1 object headers (with continuation blocks), old-style groups (symbol-table
           message -> version-1 B-tree of any depth -> symbol-table nodes -> local heap), version-1/2 dataspaces,
           fixed-point / IEEE floating-point / fixed-length string datatypes, variable-length strings through the global
           heap (attributes only), contiguous and compact data layouts (layout message versions 1-3), attributes
           (message versions 1-3).  Chunked or filtered datasets raise NotImplementedError: Keras writes small float32
           arrays contiguously, uncompressed.
  writer   the same subset, laid out the way libhdf5 1.8/1.10 with `libver='earliest'` lays files out (superblock 0,
           version-1 headers, one symbol-table node per group with a leaf K large enough for every group of the model).

CANNOT BE VALIDATED AGAINST REAL KERAS OUTPUT HERE (no h5py, no TensorFlow, no .h5 file anywhere in the image).  The writer
follows the published format specification and the reader parses what the writer produces plus the general structures listed
above; tests/test_h5lite.py checks round trips and structural invariants only.  DESIGN.md records this.
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


# ======================================================================================================================
# reader
# ======================================================================================================================
class H5Error(ValueError):
    pass


class _Node:
    """a group (children: dict) or a dataset (data: ndarray), with attrs"""

    def __init__(self):
        self.attrs = {}
        self.children = None
        self.data = None

    @property
    def is_group(self):
        return self.children is not None

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if node.children is None or part not in node.children:
                raise KeyError(path)
            node = node.children[part]
        return node

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def keys(self):
        return list(self.children.keys()) if self.children is not None else []

    def visit_datasets(self, prefix=""):
        out = []
        if self.children is None:
            return [(prefix, self.data)]
        for k, v in self.children.items():
            out += v.visit_datasets(prefix + "/" + k if prefix else k)
        return out


class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIG:
            raise H5Error("not an HDF5 file (signature)")
        ver = buf[8]
        if ver not in (0, 1):
            raise H5Error("superblock version %d is not supported (only 0 and 1: libver='earliest')" % ver)
        self.so, self.sl = buf[13], buf[14]
        if self.so != 8 or self.sl != 8:
            raise H5Error("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", buf, 16)
        p = 24 if ver == 0 else 28                      # version 1 adds indexed-storage K + reserved
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p)
        p += 32
        # root group symbol table entry
        _name_off, self.root_header = struct.unpack_from("<QQ", buf, p)

    # ---- primitives ----
    def u(self, off, fmt):
        return struct.unpack_from("<" + fmt, self.b, off)

    def read_root(self):
        return self.read_object(self.root_header)

    # ---- object headers ----
    def messages(self, addr):
        """yields (type, flags, data_offset, size) of every message of a version-1 object header"""
        b = self.b
        addr += self.base
        ver = b[addr]
        if ver != 1:
            raise H5Error("object header version %d is not supported (version-2 'OHDR' headers need libver='earliest' files)" % ver)
        n_msgs, = self.u(addr + 2, "H")
        hdr_size, = self.u(addr + 8, "I")
        blocks = [(addr + 16, hdr_size)]
        out = []
        bi = 0
        while bi < len(blocks) and len(out) < n_msgs:
            p, size = blocks[bi]
            end = p + size
            bi += 1
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize, mflags = self.u(p, "HHB")
                data = p + 8
                if mtype == 0x0010:                                   # continuation: offset, length
                    coff, clen = self.u(data, "QQ")
                    blocks.append((coff + self.base, clen))
                out.append((mtype, mflags, data, msize))
                p = data + msize
        return out

    def read_object(self, addr):
        node = _Node()
        space = dtype = layout = None
        for mtype, mflags, p, size in self.messages(addr):
            if mtype == 0x0011:                                       # symbol table: group
                btree, heap = self.u(p, "QQ")
                node.children = {}
                for name, child_addr in self.group_entries(btree, heap):
                    node.children[name] = self.read_object(child_addr)
            elif mtype == 0x0001:
                space = self.parse_dataspace(p)
            elif mtype == 0x0003:
                dtype = self.parse_datatype(p)[0]
            elif mtype == 0x0008:
                layout = self.parse_layout(p)
            elif mtype == 0x000C:
                name, value = self.parse_attribute(p)
                node.attrs[name] = value
            elif mtype == 0x000B:
                raise NotImplementedError("filtered (compressed) datasets are not supported")
        if node.children is None:
            if space is None or dtype is None or layout is None:
                raise H5Error("object at %d is neither a group nor a complete dataset" % addr)
            node.data = self.read_data(space, dtype, layout)
        return node

    # ---- groups ----
    def heap_name(self, heap_addr, off):
        heap_addr += self.base
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        data_addr, = self.u(heap_addr + 24, "Q")
        s = data_addr + self.base + off
        e = self.b.index(b"\x00", s)
        return self.b[s:e].decode("utf8")

    def group_entries(self, btree_addr, heap_addr):
        out = []
        a = btree_addr + self.base
        if self.b[a:a + 4] != b"TREE":
            raise H5Error("bad B-tree signature")
        ntype, level, used = self.u(a + 4, "BBH")
        if ntype != 0:
            raise H5Error("not a group B-tree")
        p = a + 24                                          # after signature, type, level, used, two sibling addresses
        for i in range(used):
            child, = self.u(p + 8 + i * 16, "Q")           # key i (8), child i (8), ...
            if level > 0:
                out += self.group_entries(child, heap_addr)
            else:
                out += self.snod_entries(child, heap_addr)
        return out

    def snod_entries(self, addr, heap_addr):
        a = addr + self.base
        if self.b[a:a + 4] != b"SNOD":
            raise H5Error("bad symbol-table node signature")
        n, = self.u(a + 6, "H")
        out = []
        for i in range(n):
            name_off, obj = self.u(a + 8 + 40 * i, "QQ")
            out.append((self.heap_name(heap_addr, name_off), obj))
        return out

    # ---- dataspace / datatype / layout ----
    def parse_dataspace(self, p):
        ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            q = p + 4
        else:
            raise H5Error("dataspace version %d" % ver)
        return tuple(self.u(q + 8 * i, "Q")[0] for i in range(rank))

    def parse_datatype(self, p):
        """-> (descriptor, bytes consumed); descriptor = numpy dtype, ('S', n) or ('vlen_str',)"""
        cv, b0, b1, b2 = self.b[p], self.b[p + 1], self.b[p + 2], self.b[p + 3]
        cls = cv & 0x0F
        size, = self.u(p + 4, "I")
        order = ">" if (b0 & 1) else "<"
        if cls == 0:                                        # fixed point
            signed = (b0 >> 3) & 1
            return np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), 8 + 4
        if cls == 1:                                        # floating point
            return np.dtype("%sf%d" % (order, size)), 8 + 12
        if cls == 3:                                        # fixed-length string
            return ("S", size), 8
        if cls == 9:                                        # variable length
            if (b0 & 0x0F) == 1:
                _, used = self.parse_datatype(p + 8)
                return ("vlen_str",), 8 + used
            raise NotImplementedError("variable-length sequences are not supported")
        raise NotImplementedError("datatype class %d is not supported" % cls)

    def parse_layout(self, p):
        ver = self.b[p]
        if ver == 3:
            cls = self.b[p + 1]
            if cls == 1:
                addr, size = self.u(p + 2, "QQ")
                return ("contiguous", addr, size)
            if cls == 0:
                size, = self.u(p + 2, "H")
                return ("compact", p + 4, size)
            raise NotImplementedError("chunked datasets are not supported (Keras writes contiguous float32 arrays)")
        if ver in (1, 2):
            rank, cls = self.b[p + 1], self.b[p + 2]
            q = p + 8
            if cls == 1:
                addr, = self.u(q, "Q")
                return ("contiguous", addr, None)
            if cls == 0:
                q += 4 * rank
                size, = self.u(q, "I")
                return ("compact", q + 4, size)
            raise NotImplementedError("chunked datasets are not supported")
        raise H5Error("data layout version %d" % ver)

    def decode(self, raw, shape, dtype):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dtype, np.dtype):
            a = np.frombuffer(raw, dtype, n).reshape(shape)
            return a.astype(dtype.newbyteorder("=")) if shape else a.astype(dtype.newbyteorder("="))[()]
        if dtype[0] == "S":
            a = np.frombuffer(raw, "S%d" % dtype[1], n).reshape(shape)
            return a if shape else a[()]
        if dtype[0] == "vlen_str":
            vals = []
            for i in range(n):
                length, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self.global_heap_object(gaddr, gidx)[:length])
            a = np.array(vals, dtype=object).reshape(shape)
            return a if shape else a[()]
        raise H5Error("cannot decode")

    def read_data(self, shape, dtype, layout):
        n = int(np.prod(shape)) if shape else 1
        item = dtype.itemsize if isinstance(dtype, np.dtype) else (dtype[1] if dtype[0] == "S" else 16)
        if layout[0] == "contiguous":
            if layout[1] == UNDEF:
                raw = bytes(n * item)                       # never written: fill value (zeros)
            else:
                raw = self.b[layout[1] + self.base: layout[1] + self.base + n * item]
        else:
            raw = self.b[layout[1]: layout[1] + n * item]
        return self.decode(raw, shape, dtype)

    def global_heap_object(self, addr, index):
        a = addr + self.base
        if self.b[a:a + 4] != b"GCOL":
            raise H5Error("bad global heap signature")
        size, = self.u(a + 8, "Q")
        p, end = a + 16, a + size
        while p + 16 <= end:
            idx, _ref, _res, osize = self.u(p, "HHIQ")
            if idx == index:
                return bytes(self.b[p + 16: p + 16 + osize])
            if idx == 0:
                break
            p += 16 + ((osize + 7) // 8) * 8
        raise H5Error("global heap object %d not found" % index)

    def parse_attribute(self, p):
        ver = self.b[p]
        name_size, dt_size, sp_size = self.u(p + 2, "HHH")
        if ver == 1:
            q = p + 8
            pad = lambda n: (n + 7) // 8 * 8          # noqa: E731
        elif ver in (2, 3):
            q = p + 8 + (1 if ver == 3 else 0)
            pad = lambda n: n                         # noqa: E731
        else:
            raise H5Error("attribute message version %d" % ver)
        name = bytes(self.b[q:q + name_size]).split(b"\x00")[0].decode("utf8")
        q += pad(name_size)
        dtype = self.parse_datatype(q)[0]
        q += pad(dt_size)
        shape = self.parse_dataspace(q)
        q += pad(sp_size)
        n = int(np.prod(shape)) if shape else 1
        item = dtype.itemsize if isinstance(dtype, np.dtype) else (dtype[1] if dtype[0] == "S" else 16)
        return name, self.decode(self.b[q:q + n * item], shape, dtype)


def read(path_or_bytes):
    """-> root _Node (groups: .children, datasets: .data, both: .attrs)"""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    return _Reader(buf).read_root()


# ======================================================================================================================
# writer
# ======================================================================================================================
class Group:
    def __init__(self):
        self.children = {}
        self.attrs = {}

    def group(self, name):
        g = self.children.get(name)
        if g is None:
            g = self.children[name] = Group()
        return g

    def dataset(self, name, array):
        self.children[name] = Dataset(array)
        return self.children[name]


class Dataset:
    def __init__(self, array):
        self.array = np.ascontiguousarray(array)
        self.attrs = {}


def _pad8(b):
    return b + bytes((-len(b)) % 8)


def _dt_float32():
    # class 1 version 1; LE, mantissa normalisation "implied MSB" (bits 4-5 = 2), sign bit 31; size 4;
    # bit offset 0, precision 32, exponent location 23, size 8, mantissa location 0, size 23, bias 127
    return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_float64():
    return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _dt_int(size, signed):
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def _dt_string(n):
    # class 3 version 1; null-padded, ASCII
    return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, n)


def _datatype_of(a):
    if a.dtype == np.float32:
        return _dt_float32()
    if a.dtype == np.float64:
        return _dt_float64()
    if a.dtype.kind in "iu":
        return _dt_int(a.dtype.itemsize, a.dtype.kind == "i")
    if a.dtype.kind == "S":
        return _dt_string(a.dtype.itemsize)
    raise TypeError("h5lite cannot store dtype %s" % a.dtype)


def _dataspace(shape):
    if len(shape) == 0:
        return struct.pack("<BBB5x", 1, 0, 0)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _as_attr_array(v):
    if isinstance(v, str):
        v = v.encode("utf8")
    if isinstance(v, bytes):
        return np.array(v, dtype="S%d" % max(len(v), 1))
    a = np.asarray(v)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf8")
    if a.dtype.kind == "O":
        a = np.array([x.encode("utf8") if isinstance(x, str) else x for x in a.ravel()]).reshape(a.shape)
    if a.dtype.kind == "S" and a.dtype.itemsize == 0:
        a = a.astype("S1")
    return np.ascontiguousarray(a)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_message(name, value):
    a = _as_attr_array(value)
    nm = name.encode("utf8") + b"\x00"
    dt, sp = _datatype_of(a), _dataspace(a.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + a.tobytes()
    if len(_pad8(body)) > 0xFFF8:
        raise ValueError("attribute %r is too large for a version-1 object header message (64 KB)" % name)
    return _message(0x000C, body)


class _Writer:
    LEAF_K = 32                                # one symbol-table node holds up to 2K = 64 entries
    INT_K = 16

    def __init__(self):
        self.buf = bytearray(96)               # the superblock is patched in at the end

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        off = len(self.buf)
        self.buf += data
        return off

    def object_header(self, messages):
        body = b"".join(messages)
        return self.alloc(struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)

    def write_dataset(self, d):
        a = d.array
        addr = self.alloc(a.tobytes()) if a.size else UNDEF
        msgs = [_message(0x0001, _dataspace(a.shape)),
                _message(0x0003, _datatype_of(a), flags=1),                    # constant message
                _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),             # fill value v2: late allocation, write if set, undefined
                _message(0x0008, struct.pack("<BBQQ", 3, 1, addr, a.nbytes))]   # layout v3, contiguous
        msgs += [_attr_message(k, v) for k, v in d.attrs.items()]
        return self.object_header(msgs)

    def write_group(self, g):
        if len(g.children) > 2 * self.LEAF_K:
            raise ValueError("h5lite writes one symbol-table node per group (at most %d members)" % (2 * self.LEAF_K))
        names = sorted(g.children)                                               # symbol-table nodes are sorted by name
        child_addr = {}
        for n in names:
            c = g.children[n]
            child_addr[n] = self.write_group(c) if isinstance(c, Group) else self.write_dataset(c)
        # local heap: offset 0 = "" (the B-tree's first key), then the names, then one free block
        heap = bytearray(8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\x00")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 32) + bytes(16)                            # free block: next = 1 (none), size 32
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        # one symbol-table node
        snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(names)))
        for n in names:
            c = g.children[n]
            if isinstance(c, Group):
                snod += struct.pack("<QQII", name_off[n], child_addr[n], 1, 0) + struct.pack("<QQ", *c._bt_heap)
            else:
                snod += struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0)
        snod += bytes(8 + 40 * 2 * self.LEAF_K - len(snod))
        snod_addr = self.alloc(bytes(snod))
        # B-tree node (level 0) with one child; allocated at its full size: 2K children, 2K + 1 keys
        bt = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        bt += struct.pack("<Q", 0)                                                # key 0: the empty string at heap offset 0
        if names:
            bt += struct.pack("<QQ", snod_addr, name_off[names[-1]])              # child 0, key 1 = largest name in it
        bt += bytes(24 + (2 * self.INT_K) * 8 + (2 * self.INT_K + 1) * 8 - len(bt))
        bt_addr = self.alloc(bytes(bt))
        g._bt_heap = (bt_addr, heap_addr)
        msgs = [_message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        msgs += [_attr_message(k, v) for k, v in g.attrs.items()]
        return self.object_header(msgs)

    def finish(self, root):
        root_addr = self.write_group(root)
        eof = len(self.buf)
        sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", self.LEAF_K, self.INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", *root._bt_heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write(path, root):
    data = _Writer().finish(root)
    if path is not None:
        with open(path, "wb") as f:
            f.write(data)
    return data
