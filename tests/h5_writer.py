"""Test helper: writes small HDF5 files in the layout class HighFive / libhdf5 produce by default ("earliest" format) -- superblock version 0,
version-1 object headers, symbol-table groups (version-1 B-tree -> symbol-table nodes of at most 8 entries + local heap), contiguous or compact
datasets, string attributes as variable-length strings in one global heap collection (or fixed-length strings).  It exists so the engine's HDF5
reader (raptor_b200/csrc/h5_io.cu) can be exercised on actor shapes and datatypes the one reference file (the Raptor checkpoint.h5) does not
cover: MLP actors laid out as rl_tools::save would (nn_models/{sequential,mlp}/persist.h, nn/layers/*/persist.h, nn/parameters/persist.h),
float64 / big-endian / integer data, compact layout.  Independent of the reader: it shares no code with it.

    tree = {"actor": Group({"layers": Group({"0": Group({...}, attrs={"type": "dense"})})}, attrs={...}), "example": Group({"input": array})}
    data = write(tree)
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Group:
    def __init__(self, children=None, attrs=None):
        self.children, self.attrs = dict(children or {}), dict(attrs or {})


class Dataset:
    def __init__(self, array, attrs=None, compact=False, dtype=None):
        """dtype: numpy dtype string as stored in the file ('<f4', '>f8', '<i4', ...); default little-endian of the array's own type"""
        self.array = np.asarray(array)
        self.dtype = np.dtype(dtype) if dtype else self.array.dtype.newbyteorder("<")
        self.attrs, self.compact = dict(attrs or {}), compact


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class _Writer:
    def __init__(self, fixed_strings=False):
        self.buf = bytearray(96)          # the superblock goes here at the end
        self.fixed_strings = fixed_strings
        self.gcol_at = None
        self.gcol_index = {}

    def alloc(self, data):
        at = len(self.buf)
        self.buf += _pad8(bytes(data))
        return at

    # ---- global heap with every attribute string ----------------------------------------------------------------------------------------
    def strings_of(self, node, out):
        for v in node.attrs.values():
            if v not in out:
                out.append(v)
        if isinstance(node, Group):
            for c in node.children.values():
                self.strings_of(c if isinstance(c, (Group, Dataset)) else Dataset(c), out)

    def write_gcol(self, strings):
        body = b""
        for i, s in enumerate(strings, start=1):
            raw = s.encode()
            body += struct.pack("<HHIQ", i, 1, 0, len(raw)) + _pad8(raw)
            self.gcol_index[s] = i
        free = 64
        total = 16 + len(body) + 16 + free
        body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\0" * free
        self.gcol_at = self.alloc(b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", total) + body)

    # ---- messages ----------------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def msg(mtype, body, flags=0):
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def attribute(self, name, value):
        nm = name.encode() + b"\0"
        space = bytes([1, 0, 0, 0, 0, 0, 0, 0])                                    # scalar
        if self.fixed_strings:
            raw = value.encode() + b"\0"
            dtype = bytes([0x13, 0, 0, 0]) + struct.pack("<I", len(raw))
            data = raw
        else:
            dtype = bytes([0x19, 0x01, 0, 0]) + struct.pack("<I", 16) + bytes([0x13, 0, 0, 0]) + struct.pack("<I", 1)
            data = struct.pack("<IQI", len(value.encode()), self.gcol_at, self.gcol_index[value])
        return self.msg(0xC, bytes([1, 0]) + struct.pack("<HHH", len(nm), len(dtype), len(space)) + _pad8(nm) + _pad8(dtype) + _pad8(space) + data)

    def header(self, messages):
        """version-1 object header; the last message goes into a continuation block when there are more than three (as libhdf5 does once the
        first block is full) so the reader's continuation path is used"""
        if len(messages) > 3:
            tail = b"".join(messages[3:])
            tail_at = self.alloc(tail)
            first = messages[:3] + [self.msg(0x10, struct.pack("<QQ", tail_at, len(tail)))]
        else:
            first = messages
        body = b"".join(first)
        return self.alloc(struct.pack("<BxHII4x", 1, len(messages) + (1 if len(messages) > 3 else 0), 1, len(body)) + body)

    @staticmethod
    def datatype(dt):
        big = 1 if dt.byteorder == ">" else 0
        if dt.kind == "f":
            exp_loc, exp_size, man_size, bias = {4: (23, 8, 23, 127), 8: (52, 11, 52, 1023)}[dt.itemsize]
            return bytes([0x11, 0x20 | big, dt.itemsize * 8 - 1, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, exp_loc, exp_size, 0, man_size, bias)
        if dt.kind in "iu":
            return bytes([0x10, big | (8 if dt.kind == "i" else 0), 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
        raise ValueError(dt)

    def dataset(self, ds):
        raw = ds.array.astype(ds.dtype).tobytes()
        dims = ds.array.shape
        space = bytes([1, len(dims), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", d) for d in dims)
        if ds.compact:
            layout = bytes([3, 0]) + struct.pack("<H", len(raw)) + raw
        else:
            layout = bytes([3, 1]) + struct.pack("<QQ", self.alloc(raw) if raw else UNDEF, len(raw))
        msgs = [self.msg(0x1, space), self.msg(0x3, self.datatype(ds.dtype), flags=1), self.msg(0x8, layout)]
        msgs += [self.attribute(k, v) for k, v in ds.attrs.items()]
        return self.header(msgs)

    def group(self, g):
        entries = []
        for name in sorted(g.children):
            c = g.children[name]
            if not isinstance(c, (Group, Dataset)):
                c = Dataset(c)
            entries.append((name, self.group(c) if isinstance(c, Group) else self.dataset(c)))
        # local heap: offset 0 holds the empty string
        seg, offsets = bytearray(8), {}
        for name, _ in entries:
            offsets[name] = len(seg)
            seg += _pad8(name.encode() + b"\0")
        seg += b"\0" * 16
        seg_at = self.alloc(seg)
        heap_at = self.alloc(b"HEAP" + bytes(4) + struct.pack("<QQQ", len(seg), UNDEF, seg_at))
        # symbol-table nodes of at most 8 symbols under one B-tree node
        keys, children = [0], []
        for i in range(0, len(entries), 8):
            chunk = entries[i:i + 8]
            snod = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(chunk))
            for name, addr in chunk:
                snod += struct.pack("<QQII16x", offsets[name], addr, 0, 0)
            children.append(self.alloc(snod))
            keys.append(offsets[chunk[-1][0]])
        tree = b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", len(children), UNDEF, UNDEF)
        for k, child in zip(keys, children):
            tree += struct.pack("<QQ", k, child)
        tree += struct.pack("<Q", keys[-1])
        tree_at = self.alloc(tree)
        msgs = [self.msg(0x11, struct.pack("<QQ", tree_at, heap_at))] + [self.attribute(k, v) for k, v in g.attrs.items()]
        g._tree_heap = (tree_at, heap_at)
        return self.header(msgs)


def write(tree, fixed_strings=False, userblock=0, superblock_version=0):
    """tree: Group or dict of the root's children.  userblock: 0 or a power of two >= 512 -- the superblock then sits at that offset and every
    address is relative to it (base address), as h5py's `userblock_size` produces.  superblock_version 1 adds the indexed-storage K field."""
    root = tree if isinstance(tree, Group) else Group(tree)
    w = _Writer(fixed_strings)
    sb_len = 96 + (4 if superblock_version == 1 else 0)
    w.buf = bytearray(sb_len + (-sb_len % 8))
    strings = []
    w.strings_of(root, strings)
    if not fixed_strings:
        w.write_gcol(strings)
    root_at = w.group(root)
    tree_at, heap_at = root._tree_heap
    sb = b"\x89HDF\r\n\x1a\n" + bytes([superblock_version, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    if superblock_version == 1:
        sb += struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, len(w.buf), UNDEF)
    sb += struct.pack("<QQII", 0, root_at, 1, 0) + struct.pack("<QQ", tree_at, heap_at)
    assert len(sb) == sb_len
    w.buf[:sb_len] = sb
    return bytes(userblock) + bytes(w.buf)
