"""
Minimal pure-Python HDF5 writer/reader for Keras-2.0.0 `save_weights` files (RUN.h5).

There is no h5py / libhdf5 in this environment, but the saved-weights layout is part of the
drop-in contract (utils/model_utils.py:138 -> model.save_weights; cl_vrnn/model.py:281
model.load_weights).  Layout written [K2-recall of keras.engine.topology.save_weights_to_hdf5_group]:

  /            attrs: layer_names (S-array, every layer of model.layers, weight-less ones too),
                      backend, keras_version (scalar fixed strings)
  /<layer>     attrs: weight_names (S-array; an empty float64 (0,) array for weight-less layers,
                      which is what h5py stores for `[]`)
  /<layer>/<layer>/<weight>:0   float32 contiguous dataset   (weight names contain '/')

File structure: superblock v0, old-style groups (symbol table message -> v1 B-tree -> SNOD + local
heap), v1 object headers, v1 dataspace / datatype messages, v3 contiguous layout -- i.e. what
libhdf5 1.8 with libver='earliest' (h5py's default) produces, with the group-leaf K raised in the
superblock so every group fits one symbol node.  NOT validated against real h5py here (unavailable);
the reader below round-trips it and also follows continuation blocks / multi-node B-trees so files
written by real Keras should load.
"""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 32, 16


def _pad8(b):
    return b + b"\x00" * ((-len(b)) % 8)


# ------------------------------------------------------------------------------ writer
class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # superblock placeholder

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- messages
    @staticmethod
    def _msg(mtype, data, flags=0):
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    @staticmethod
    def _dataspace(shape):
        if shape is None:  # scalar
            return struct.pack("<BBB5x", 1, 0, 0)
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)

    @staticmethod
    def _dtype_float(size):
        if size == 4:
            return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)

    @staticmethod
    def _dtype_string(n):
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, n)  # fixed length, null-padded, ASCII

    def _attr(self, name, value):
        name_b = name.encode() + b"\x00"
        if isinstance(value, bytes):  # scalar fixed string
            dt, ds, data = self._dtype_string(max(len(value), 1)), self._dataspace(None), value or b"\x00"
        else:
            value = list(value)
            if len(value) == 0:       # h5py stores [] as an empty float64 array
                dt, ds, data = self._dtype_float(8), self._dataspace((0,)), b""
            else:
                n = max(len(v) for v in value)
                dt, ds = self._dtype_string(n), self._dataspace((len(value),))
                data = b"".join(v.ljust(n, b"\x00") for v in value)
        body = struct.pack("<BxHHH", 1, len(name_b), len(dt), len(ds)) + _pad8(name_b) + _pad8(dt) + _pad8(ds) + data
        return self._msg(0x000C, body)

    def _object_header(self, msgs):
        body = b"".join(msgs)
        return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body

    # ---- objects
    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        daddr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [self._msg(0x0001, self._dataspace(arr.shape)), self._msg(0x0003, self._dtype_float(4), 1),
                self._msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),
                self._msg(0x0008, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes))]
        return self.alloc(self._object_header(msgs))

    def group(self, children, attrs):
        """children: dict name -> object header address.  Returns (header address, btree, heap)."""
        names = sorted(children, key=lambda s: s.encode())
        if len(names) > 2 * LEAF_K:
            raise ValueError("too many entries in one group for this writer")
        heap_data = bytearray(8)  # offset 0 = empty string
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b"\x00")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # one free block: next = 1 (none), size 16
        data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
        snod += b"\x00" * (40 * (2 * LEAF_K - len(names)))
        snod_addr = self.alloc(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        body = struct.pack("<Q", 0)
        if names:
            body += struct.pack("<QQ", snod_addr, offs[names[-1]])
        body += b"\x00" * ((2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(body))
        btree_addr = self.alloc(tree + body)
        msgs = [self._msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [self._attr(k, v) for k, v in attrs.items()]
        return self.alloc(self._object_header(msgs)), btree_addr, heap_addr

    def finish(self, root):
        hdr, btree, heap = root
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def save_keras_weights(path, layers, backend=b"tensorflow", keras_version=b"2.0.0"):
    """layers: list of (layer_name, [weight names like 'hW/kernel:0'], [arrays]) in model.layers
    order (weight-less layers have empty lists)."""
    w = _Writer()
    top = {}
    for lname, wnames, arrays in layers:
        sub = {}
        for wn, arr in zip(wnames, arrays):
            parts = wn.split("/")
            node = sub
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = np.asarray(arr)

        def build(node):
            ch = {}
            for k, v in node.items():
                ch[k] = build(v)[0] if isinstance(v, dict) else w.dataset(v)
            return w.group(ch, {})
        ch = {}
        for k, v in sub.items():
            ch[k] = build(v)[0] if isinstance(v, dict) else w.dataset(v)
        top[lname] = w.group(ch, {"weight_names": [n.encode() for n in wnames]})[0]
    root = w.group(top, {"layer_names": [l[0].encode() for l in layers], "backend": backend,
                         "keras_version": keras_version})
    with open(path, "wb") as f:
        f.write(w.finish(root))


# ------------------------------------------------------------------------------ reader
class _Reader:
    def __init__(self, data):
        self.d = data
        if data[:8] != SIG:
            raise ValueError("not an HDF5 file")
        ver = data[8]
        if ver not in (0, 1):
            raise ValueError("superblock version %d not supported by this minimal reader" % ver)
        if data[13] != 8 or data[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", data, 16)
        pos = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", data, pos)[0]
        self.root_hdr = struct.unpack_from("<Q", data, pos + 32 + 8)[0]

    def messages(self, addr):
        d = self.d
        ver, nmsg, _, size = struct.unpack_from("<BxHII", d, addr)
        if ver != 1:
            raise ValueError("object header v%d not supported" % ver)
        out, blocks = [], [(addr + 16, size)]
        while blocks:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", d, pos)
                body = d[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, body))
        return out

    def _dtype(self, b):
        cls, ver = b[0] & 0x0F, b[0] >> 4
        size = struct.unpack_from("<I", b, 4)[0]
        if cls == 1:
            return ("f", size, 8 + 12)
        if cls == 3:
            return ("S", size, 8)
        if cls == 0:
            return ("i", size, 8 + 4)
        raise ValueError("datatype class %d not supported" % cls)

    def _dataspace(self, b):
        ver, rank, flags = b[0], b[1], b[2]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", b, off + 8 * i)[0] for i in range(rank)), rank

    def attrs(self, addr):
        out = {}
        for mtype, b in self.messages(addr):
            if mtype != 0x000C:
                continue
            ver = b[0]
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
            pos = 8
            rnd = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
            name = b[pos:pos + nsz].split(b"\x00")[0].decode(); pos += rnd(nsz)
            kind, size, _ = self._dtype(b[pos:pos + tsz]); pos += rnd(tsz)
            shape, rank = self._dataspace(b[pos:pos + ssz]); pos += rnd(ssz)
            n = int(np.prod(shape)) if rank else 1
            raw = b[pos:pos + n * size]
            if kind == "S":
                vals = [raw[i * size:(i + 1) * size].rstrip(b"\x00") for i in range(n)]
                out[name] = vals if rank else vals[0]
            else:
                out[name] = np.frombuffer(raw, dtype="<%s%d" % (kind, size)).reshape(shape)
        return out

    def children(self, addr):
        for mtype, b in self.messages(addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", b)
                hsize, _, hdata = struct.unpack_from("<QQQ", self.d, heap + 8)
                out = {}
                self._walk(btree, hdata, out)
                return out
        return None

    def _walk(self, node, hdata, out):
        d = self.d
        if d[node:node + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        level, used = struct.unpack_from("<BH", d, node + 5)
        pos = node + 24
        for i in range(used):
            child = struct.unpack_from("<Q", d, pos + 8)[0]
            pos += 16
            if level > 0:
                self._walk(child, hdata, out)
            else:
                if d[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol node")
                nsym = struct.unpack_from("<H", d, child + 6)[0]
                for j in range(nsym):
                    noff, hdr = struct.unpack_from("<QQ", d, child + 8 + 40 * j)
                    name = d[hdata + noff:d.index(b"\x00", hdata + noff)].decode()
                    out[name] = hdr

    def dataset(self, addr):
        shape = kind = size = None
        layout = None
        for mtype, b in self.messages(addr):
            if mtype == 0x0001:
                shape, _ = self._dataspace(b)
            elif mtype == 0x0003:
                kind, size, _ = self._dtype(b)
            elif mtype == 0x0008:
                layout = b
        if layout is None or shape is None:
            return None
        n = int(np.prod(shape)) if len(shape) else 1
        if layout[0] != 3:
            raise ValueError("data layout v%d not supported" % layout[0])
        if layout[1] == 1:
            daddr, dsize = struct.unpack_from("<QQ", layout, 2)
            raw = self.d[daddr:daddr + n * size] if n else b""
        elif layout[1] == 0:
            csize = struct.unpack_from("<H", layout, 2)[0]
            raw = layout[4:4 + csize]
        else:
            raise ValueError("chunked datasets not supported by this minimal reader")
        return np.frombuffer(raw, dtype="<%s%d" % (kind, size)).reshape(shape).copy()

    def find(self, addr, path):
        for p in path.split("/"):
            ch = self.children(addr)
            addr = ch[p]
        return addr


def load_keras_weights(path):
    """-> list of (layer_name, [arrays in weight_names order]) in layer_names order."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    root = r.root_hdr
    names = [n.decode() for n in r.attrs(root)["layer_names"]]
    top = r.children(root)
    out = []
    for lname in names:
        g = top[lname]
        wn = r.attrs(g).get("weight_names", [])
        wn = [] if isinstance(wn, np.ndarray) else [n.decode() for n in wn]
        out.append((lname, [r.dataset(r.find(g, n)) for n in wn]))
    return out
