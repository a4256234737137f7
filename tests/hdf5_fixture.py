"""Test helper: lay out a small HDF5 file the way h5py's defaults (libver "earliest") do.

Structures follow the HDF5 File Format Specification: version-0 superblock, groups as symbol
tables (local heap + version-1 B-tree over symbol nodes, leaf K = 4 / internal K = 16 as the
library's defaults), version-1 object headers, and datasets with contiguous, compact or
chunked (shuffle + deflate) storage.  It exists so that joshupscale_b200/hdf5_lite.py and the
importer's `.weights.h5` route can be exercised without h5py; it is not a general writer.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Tuple, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4
INTERNAL_K = 16

Tree = Dict[str, Union["Tree", np.ndarray, Tuple[np.ndarray, str]]]


class _Image:
    def __init__(self, userblock: int, new_style: bool = False):
        self.buf = bytearray(userblock)
        self.base = userblock
        self.new_style = new_style

    def alloc(self, data: bytes) -> int:
        """Append 8-byte aligned; returns the address relative to the base address."""
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf) - self.base
        self.buf += data
        return addr

    def patch(self, addr: int, data: bytes) -> None:
        self.buf[self.base + addr:self.base + addr + len(data)] = data


def _message(mtype: int, body: bytes) -> Tuple[int, bytes]:
    return mtype, body


def _object_header(img: "_Image", messages: List[Tuple[int, bytes]]) -> int:
    """Writes the header, returns its address.  Version 1 by default; with `img.new_style` a
    version-2 header ("OHDR") whose last message sits in a continuation chunk ("OCHK").  The
    version-2 checksums are written as zero: readers that verify them (libhdf5) would refuse
    the file, hdf5_lite does not verify them."""
    if not img.new_style:
        body = b""
        for mtype, data in messages:
            data = data + b"\0" * (-len(data) % 8)
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        return img.alloc(struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)
    enc = [struct.pack("<BHB", mtype, len(data), 0) + data for mtype, data in messages]
    tail = b""
    if len(enc) > 1:
        tail_addr = img.alloc(b"OCHK" + enc[-1] + b"\0" * 4)
        enc[-1] = struct.pack("<BHB", 0x10, 16, 0) + struct.pack("<QQ", tail_addr, 4 + len(enc[-1]) + 4)
    body = b"".join(enc) + tail
    return img.alloc(b"OHDR" + struct.pack("<BBI", 2, 0x02, len(body)) + body + b"\0" * 4)


def _datatype(dtype: np.dtype) -> bytes:
    dtype = np.dtype(dtype)
    big = 1 if dtype.byteorder == ">" else 0
    if dtype.kind == "f":
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dtype.itemsize]
        sign, eloc, esize, mloc, msize, bias = spec
        return (struct.pack("<BBBBI", 0x11, 0x20 | big, sign, 0, dtype.itemsize) +
                struct.pack("<HHBBBBI", 0, 8 * dtype.itemsize, eloc, esize, mloc, msize, bias))
    signed = 0x08 if dtype.kind == "i" else 0
    return (struct.pack("<BBBBI", 0x10, signed | big, 0, 0, dtype.itemsize) +
            struct.pack("<HH", 0, 8 * dtype.itemsize))


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _write_dataset(img: _Image, arr: np.ndarray, layout: str) -> int:
    arr = np.asarray(arr, order="C")  # (ascontiguousarray would turn a scalar into shape (1,))
    msgs = [_message(0x1, _dataspace(arr.shape)), _message(0x3, _datatype(arr.dtype))]
    raw = arr.tobytes()
    if layout == "compact":
        msgs.append(_message(0x8, struct.pack("<BBH", 3, 0, len(raw)) + raw))
    elif layout == "contiguous":
        addr = img.alloc(raw) if raw else UNDEF
        msgs.append(_message(0x8, struct.pack("<BBQQ", 3, 1, addr, len(raw))))
    elif layout == "chunked":
        rank = arr.ndim
        chunk = tuple(max(1, (s + 1) // 2) for s in arr.shape)  # 2 chunks per axis, ragged edges
        item = arr.dtype.itemsize
        entries = []
        for idx in np.ndindex(*[-(-s // c) for s, c in zip(arr.shape, chunk)]):
            offs = tuple(i * c for i, c in zip(idx, chunk))
            block = np.zeros(chunk, arr.dtype)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, arr.shape))
            block[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
            body = block.tobytes()
            n = len(body) // item
            body = np.frombuffer(body, np.uint8).reshape(n, item).T.tobytes()  # shuffle
            body = zlib.compress(body, 6)
            entries.append((offs, len(body), img.alloc(body)))
        node = struct.pack("<4sBBHQQ", b"TREE", 1, 0, len(entries), UNDEF, UNDEF)
        for offs, nbytes, addr in entries:
            node += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs + (0,))
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape + (0,))
        btree = img.alloc(node)
        msgs.append(_message(0x8, struct.pack("<BBBQ", 3, 2, rank + 1, btree) +
                             b"".join(struct.pack("<I", c) for c in chunk + (item,))))
        filters = struct.pack("<BB6x", 1, 2)
        filters += struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I", item) + b"\0" * 4
        filters += struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", 6) + b"\0" * 4
        msgs.append(_message(0xB, filters))
    else:
        raise ValueError(layout)
    return _object_header(img, msgs)


def _write_group(img: _Image, tree: Tree) -> Tuple[int, int, int]:
    """Returns (object header address, B-tree address, heap address)."""
    children: List[Tuple[str, int, Tuple[int, int]]] = []
    for name in sorted(tree):
        node = tree[name]
        if isinstance(node, dict):
            hdr, bt, hp = _write_group(img, node)
            children.append((name, hdr, (bt, hp)))
        else:
            arr, layout = node if isinstance(node, tuple) else (node, "contiguous")
            children.append((name, _write_dataset(img, np.asarray(arr), layout), (0, 0)))
    if img.new_style:
        # compact new-style group: a link-info message (no fractal heap) and one link message per child
        msgs = [_message(0x2, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF))]
        for name, hdr, _ in children:
            enc = name.encode()
            msgs.append(_message(0x6, struct.pack("<BBB", 1, 0, len(enc)) + enc + struct.pack("<Q", hdr)))
        return _object_header(img, msgs), 0, 0
    # local heap: offset 0 holds the empty string
    heap = bytearray(8)
    offsets = {}
    for name, _, _ in children:
        offsets[name] = len(heap)
        enc = name.encode() + b"\0"
        heap += enc + b"\0" * (-len(enc) % 8)
    heap_data = img.alloc(bytes(heap))
    heap_addr = img.alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), UNDEF, heap_data))
    # symbol nodes of at most 2*LEAF_K entries, B-tree levels of at most 2*INTERNAL_K children
    level: List[Tuple[int, int]] = []  # (node address, heap offset of its largest name)
    for i in range(0, max(len(children), 1), 2 * LEAF_K):
        part = children[i:i + 2 * LEAF_K]
        snod = struct.pack("<4sBxH", b"SNOD", 1, len(part))
        for name, hdr, (bt, hp) in part:
            cache = 1 if bt else 0
            snod += struct.pack("<QQI4xQQ", offsets[name], hdr, cache, bt, hp)
        level.append((img.alloc(snod), offsets[part[-1][0]] if part else 0))
    depth = 0
    while True:
        nxt: List[Tuple[int, int]] = []
        for i in range(0, len(level), 2 * INTERNAL_K):
            part = level[i:i + 2 * INTERNAL_K]
            node = struct.pack("<4sBBHQQ", b"TREE", 0, depth, len(part), UNDEF, UNDEF) + struct.pack("<Q", 0)
            for addr, key in part:
                node += struct.pack("<QQ", addr, key)
            nxt.append((img.alloc(node), part[-1][1]))
        level = nxt
        depth += 1
        if len(level) == 1:
            break
    btree = level[0][0]
    hdr = _object_header(img, [_message(0x11, struct.pack("<QQ", btree, heap_addr))])
    return hdr, btree, heap_addr


def write_hdf5(path: str, tree: Tree, userblock: int = 0, new_style: bool = False) -> None:
    """`tree`: nested dicts; leaves are arrays or (array, "contiguous" | "compact" | "chunked").
    `new_style`: version-2 superblock and object headers, groups as link messages (what h5py writes
    with libver="latest" for small groups)."""
    img = _Image(userblock, new_style)
    sb_at = img.alloc(b"\0" * 96)
    root_hdr, btree, heap = _write_group(img, tree)
    eof = len(img.buf) - img.base
    if new_style:
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0)
        sb += struct.pack("<QQQQI", userblock, UNDEF, eof, root_hdr, 0)
        img.patch(sb_at, sb)
        with open(path, "wb") as f:
            f.write(bytes(img.buf))
        return
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, eof, UNDEF)  # addresses are relative to the base
    sb += struct.pack("<QQI4xQQ", 0, root_hdr, 1, btree, heap)
    img.patch(sb_at, sb)
    with open(path, "wb") as f:
        f.write(bytes(img.buf))
