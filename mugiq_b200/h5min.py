"""Self-contained HDF5 writer (and a small reader) for the loop files of Loop_Mugiq::writeLoopsHDF5_Mom
(/root/reference/lib/loop_mugiq.cpp:530-656): group tree /mom_%+d_%+d_%+d/<disp tag>/<GammaName>/ with one dataset "loop"
of shape [T][2] (real, imag) in native little-endian double or float — the on-disk contract downstream analysis reads
(SURVEY §8f rank 1).  No HDF5 library exists in this image (no libhdf5, h5py or PyTables), so the file is produced
directly from the HDF5 File Format Specification, in the oldest and most widely readable dialect, which is also what
`H5Fcreate` with default properties writes:

  superblock version 0 (8-byte offsets and lengths) · "old style" groups = symbol-table message -> v1 B-tree ("TREE",
  node type 0) -> symbol-table nodes ("SNOD") + local heap ("HEAP") for the link names · version-1 object headers ·
  dataset = dataspace v1 + datatype v1 (IEEE little-endian float) + fill-value v2 (undefined) + layout v3 contiguous.

Every group gets ONE symbol-table node: the file-level "group leaf node K" of the superblock (H5Pset_sym_k) is chosen so
that the largest group fits, and entries are sorted by link name as the B-tree requires.  Links are hard links, every
object is referenced once, nothing is compressed, chunked, shared or timestamped, so the output is byte-for-byte
deterministic.

What can be verified here: `read()` below is an independent reader of the same dialect that is ALSO checked against a
real HDF5 file written by the HDF5 library (a MATLAB 7.3 file shipped with scipy's test data, tests/test_h5min.py);
writer -> reader round trips are bit-exact.  What cannot: libhdf5 itself reading these files (it is absent) — the quirk
that would matter most is handled (a local heap without free blocks stores H5HL_FREE_NULL = 1, not the undefined
address, as its free-list head).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"
INTERNAL_K = 16  # library default (H5B node of 2K children); one child is ever used


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class _Node:
    def __init__(self):
        self.children = {}  # name -> _Node | np.ndarray


def _tree(datasets):
    root = _Node()
    for path, arr in datasets.items():
        parts = [p for p in path.split("/") if p]
        if not parts:
            raise ValueError("empty dataset path")
        node = root
        for p in parts[:-1]:
            nxt = node.children.setdefault(p, _Node())
            if not isinstance(nxt, _Node):
                raise ValueError(f"{path}: {p} is a dataset")
            node = nxt
        if parts[-1] in node.children:
            raise ValueError(f"{path}: duplicate")
        a = np.ascontiguousarray(arr)
        if a.dtype not in (np.float64, np.float32):
            raise TypeError(f"{path}: only float32 / float64 datasets are supported, got {a.dtype}")
        node.children[parts[-1]] = a
    return root


def _max_entries(node):
    m = len(node.children)
    for c in node.children.values():
        if isinstance(c, _Node):
            m = max(m, _max_entries(c))
    return m


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages):
    body = b"".join(messages)
    # version 1, reserved, #messages, reference count 1, size of the message block; the prefix is padded to 16 bytes
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _datatype_float(itemsize):
    if itemsize == 8:
        sign, prec, eloc, esize, msize, bias = 63, 64, 52, 11, 52, 1023
    else:
        sign, prec, eloc, esize, msize, bias = 31, 32, 23, 8, 23, 127
    # class 1 (floating point), version 1; bit field: little endian, mantissa normalisation 2 (msb implied), sign location
    return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, itemsize) + struct.pack("<HHBBBBI", 0, prec, eloc, esize, 0, msize, bias)


class _Writer:
    def __init__(self, leaf_k):
        self.leaf_k = leaf_k
        self.buf = bytearray(96)  # superblock written last
        self.raw_offset = {}      # dataset path -> byte offset of its raw data in the file

    def alloc(self, data):
        off = len(self.buf)
        assert off % 8 == 0
        self.buf += _pad8(bytes(data))
        return off

    def dataset(self, arr, path=None):
        raw = self.alloc(arr.tobytes()) if arr.size else UNDEF
        if path is not None:
            self.raw_offset[path] = raw
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)  # v2: allocate late, write fill if set, fill value undefined
        layout = struct.pack("<BBQQ", 3, 1, raw, arr.nbytes)
        # message flags as the library sets them (bit 0 = constant): dataspace 0, datatype / fill value / layout 1
        return self.alloc(_object_header([_message(0x0001, space, 0), _message(0x0003, _datatype_float(arr.itemsize), 1),
                                          _message(0x0005, fill, 1), _message(0x0008, layout, 1)]))

    def group(self, node, prefix=""):
        """Writes the objects below `node`, then its heap, symbol-table node, B-tree and object header.
        Returns (object header address, B-tree address, heap address)."""
        names = sorted(node.children, key=lambda s: s.encode())
        entries = []
        for name in names:
            child = node.children[name]
            if isinstance(child, _Node):
                oh, bt, hp = self.group(child, prefix + "/" + name)
                entries.append((name, oh, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name, self.dataset(child, prefix + "/" + name), 0, b"\0" * 16))
        # local heap: "" at offset 0, then the names; a 16-byte free block closes the segment
        heap = bytearray(_pad8(b"\0"))
        offs = []
        for name, *_ in entries:
            offs.append(len(heap))
            heap += _pad8(name.encode() + b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)  # next free block: none (H5HL_FREE_NULL), size of this block
        heap_data = self.alloc(heap)
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        # symbol-table node: 2 * leaf_k slots
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)))
        for (name, oh, cache, scratch), o in zip(entries, offs):
            snod += struct.pack("<QQII", o, oh, cache, 0) + scratch
        snod += b"\0" * (40 * (2 * self.leaf_k - len(entries)))
        snod_addr = self.alloc(snod)
        # B-tree leaf (level 0) with one child: keys are heap offsets of the smallest ("") and largest name
        nchild = 1 if entries else 0
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, nchild, UNDEF, UNDEF))
        slots = struct.pack("<Q", 0)
        if entries:
            slots += struct.pack("<QQ", snod_addr, offs[-1])
        tree += slots + b"\0" * ((2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(slots))
        tree_addr = self.alloc(tree)
        oh_addr = self.alloc(_object_header([_message(0x0011, struct.pack("<QQ", tree_addr, heap_addr), 1)]))
        return oh_addr, tree_addr, heap_addr

    def finish(self, root):
        oh, bt, hp = self.group(root)
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def dumps(datasets):
    """{"/group/.../name": float32/float64 ndarray} -> bytes of an HDF5 file."""
    root = _tree(datasets)
    leaf_k = max(4, (_max_entries(root) + 1) // 2)  # 4 is the library default
    if leaf_k > 0xFFFF:
        raise ValueError("too many links in one group")
    return _Writer(leaf_k).finish(root)


def write(filename, datasets):
    blob = dumps(datasets)
    with open(filename, "wb") as fh:
        fh.write(blob)
    return len(blob)


# ---- several writers, one file: what the reference does with MPI-IO and hyperslabs ---------------------------------------
# writeLoopsHDF5_Mom opens the file collectively (H5Pset_fapl_mpio, /root/reference/lib/loop_mugiq.cpp:571-572), creates every
# [totT][2] dataset on all ranks and lets each "time process" write its rows [tCoord*locT, +locT) (:561-565, :624).  With
# contiguous datasets that is: ONE rank lays the file out (metadata + zero-filled raw data), then every rank writes its rows
# at (raw data offset of the dataset) + row0 * (bytes per row) - positional writes into disjoint byte ranges of a shared file.
def skeleton(filename, shapes, dtype=np.float64):
    """Writes the file with every dataset of `shapes` ({path: shape}) zero-filled and returns {path: raw data offset}.
    The layout depends on the paths and shapes only: every rank can compute the same offsets with `offsets()`."""
    w, root = _layout(shapes, dtype)
    blob = w.finish(root)
    with open(filename, "wb") as fh:
        fh.write(blob)
    return dict(w.raw_offset)


def _layout(shapes, dtype):
    datasets = {p: np.zeros(shape, dtype=dtype) for p, shape in shapes.items()}
    root = _tree(datasets)
    leaf_k = max(4, (_max_entries(root) + 1) // 2)
    if leaf_k > 0xFFFF:
        raise ValueError("too many links in one group")
    return _Writer(leaf_k), root


def offsets(shapes, dtype=np.float64):
    """{path: raw data offset} of the file `skeleton` writes for these shapes, without writing anything."""
    w, root = _layout(shapes, dtype)
    w.finish(root)
    return dict(w.raw_offset)


def write_rows(filename, offset, row0, rows):
    """Rows [row0, row0 + len(rows)) of the dataset whose raw data start at byte `offset` of an existing file
    (the hyperslab [tCoord*locT, +locT) of lib/loop_mugiq.cpp:624).  `rows`: C-contiguous array, first axis = rows."""
    import os
    rows = np.ascontiguousarray(rows)
    row_bytes = rows.nbytes // max(rows.shape[0], 1)
    fd = os.open(filename, os.O_WRONLY)
    try:
        os.pwrite(fd, rows.tobytes(), offset + row0 * row_bytes)
    finally:
        os.close(fd)


# ---- reader (old-style groups, v1 object headers, contiguous / compact float and integer datasets) ---------------------
class _Reader:
    def __init__(self, blob):
        self.b = blob
        self.base = None
        for off in [0] + [512 << i for i in range(12)]:  # a user block (MATLAB: 512 bytes) may precede the superblock
            if blob[off:off + 8] == SIG:
                self.base = off
                break
        if self.base is None:
            raise ValueError("not an HDF5 file")
        sb = blob[self.base:]
        if sb[8] != 0 or sb[13] != 8 or sb[14] != 8:
            raise ValueError("only superblock version 0 with 8-byte offsets / lengths is supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", sb, 16)
        self.root_oh = struct.unpack_from("<Q", sb, 56 + 8)[0]

    def at(self, addr, n):
        a = self.base + addr
        return self.b[a:a + n]

    def messages(self, addr):
        ver, _, nmsg, _, size = struct.unpack("<BBHII", self.at(addr, 12))
        if ver != 1:
            raise ValueError(f"object header version {ver} at {addr}")
        out, blocks = [], [(addr + 16, size)]
        while blocks:
            pos, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack("<HHB", self.at(pos, 5))
                data = self.at(pos + 8, msize)
                pos += 8 + msize
                left -= 8 + msize
                if mtype == 0x0010:  # continuation
                    blocks.append(struct.unpack("<QQ", data[:16]))
                out.append((mtype, data))
        return out

    def heap_name(self, heap_addr, off):
        h = self.at(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise ValueError("bad local heap")
        data_addr = struct.unpack_from("<Q", h, 24)[0]
        a = self.base + data_addr + off
        return self.b[a:self.b.index(b"\0", a)].decode()

    def links(self, tree_addr, heap_addr):
        node = self.at(tree_addr, 24)
        if node[:4] != b"TREE" or node[4] != 0:
            raise ValueError("bad group B-tree node")
        level, used = node[5], struct.unpack_from("<H", node, 6)[0]
        out = []
        for i in range(used):
            child = struct.unpack("<Q", self.at(tree_addr + 24 + 8 + 16 * i, 8))[0]
            if level > 0:
                out += self.links(child, heap_addr)
                continue
            sn = self.at(child, 8)
            if sn[:4] != b"SNOD":
                raise ValueError("bad symbol-table node")
            for k in range(struct.unpack_from("<H", sn, 6)[0]):
                name_off, oh = struct.unpack("<QQ", self.at(child + 8 + 40 * k, 16))
                out.append((self.heap_name(heap_addr, name_off), oh))
        return out

    def walk(self, oh_addr, prefix, out):
        msgs = dict(self.messages(oh_addr))
        if 0x0011 in msgs:  # group
            tree, heap = struct.unpack("<QQ", msgs[0x0011][:16])
            names = []
            for name, child in self.links(tree, heap):
                names.append(name)
                self.walk(child, prefix + "/" + name, out)
            if names != sorted(names, key=lambda s: s.encode()):
                raise ValueError(f"{prefix or '/'}: links are not sorted")
            return
        if 0x0008 not in msgs:
            return  # named datatype or something else this reader does not need
        sp = msgs[0x0001]
        rank = sp[1]
        dims = struct.unpack_from(f"<{rank}Q", sp, 8 if sp[0] == 1 else 4)
        dt = msgs[0x0003]
        cls, size = dt[0] & 0x0F, struct.unpack_from("<I", dt, 4)[0]
        if dt[1] & 1:
            raise ValueError("big-endian data")
        if cls == 1:
            np_dt = {4: np.float32, 8: np.float64}[size]
        elif cls == 0:
            np_dt = np.dtype(("i" if dt[1] & 8 else "u") + str(size))
        else:
            out[prefix or "/"] = None  # e.g. MATLAB's references / strings
            return
        lay = msgs[0x0008]
        n = int(np.prod(dims)) if rank else 1
        if lay[0] == 3 and lay[1] == 1:  # contiguous
            addr, nbytes = struct.unpack_from("<QQ", lay, 2)
            raw = b"" if addr == UNDEF else self.at(addr, nbytes)
        elif lay[0] == 3 and lay[1] == 0:  # compact
            nbytes = struct.unpack_from("<H", lay, 2)[0]
            raw = lay[4:4 + nbytes]
        elif lay[0] in (1, 2) and lay[2] == 1:  # versions 1 and 2, contiguous: dimensions then the element size
            nd = lay[1]
            addr = struct.unpack_from("<Q", lay, 8)[0]
            nbytes = int(np.prod(struct.unpack_from(f"<{nd}I", lay, 16)))
            raw = self.at(addr, nbytes)
        else:
            out[prefix or "/"] = None  # chunked
            return
        out[prefix or "/"] = np.frombuffer(raw, dtype=np_dt, count=n).reshape(dims).copy()


def loads(blob):
    """bytes of an HDF5 file in the dialect above -> {"/path": ndarray (None for datasets this reader cannot decode)}."""
    r = _Reader(blob)
    out = {}
    r.walk(r.root_oh, "", out)
    return out


def read(filename):
    with open(filename, "rb") as fh:
        return loads(fh.read())
