"""Minimal ASDF reader / writer for the hot path's files.

The reference opens every file of the path with ``asdf.open`` (`L1_to_L2/gen_cal_image.py:112,172,213,531,536,560,632`,
`L1_to_L2/oututils.py:43`, `utils/ipc_linearity.py:170,267,324,380`, `utils/fitting.py:201,207`,
`utils/flatutils.py:47,63,71`) and only ever does three things with the result: index the YAML tree with string keys,
read whole ndarray leaves (or leading-axis slices of them), and - for the output - write a tree of dicts / arrays.
This module does exactly that, from the published ASDF standard (1.x), with PyYAML + NumPy only:

* file = ``#ASDF`` header comments, one YAML 1.1 document (tags kept, not interpreted), then binary blocks:
  ``\\xd3BLK`` | header size u16 | flags u32 | compression 4 bytes | allocated u64 | used u64 | data u64 | MD5 16 bytes
  (all big-endian) | payload; an optional ``#ASDF BLOCK INDEX`` document at the end;
* ``!core/ndarray`` leaves (`source`, `datatype`, `byteorder`, `shape`, optional `offset` / `strides`) become
  :class:`BlockArray` objects: ``np.asarray(leaf)`` / ``leaf[...]`` memory-map uncompressed blocks (zero copy) and
  inflate ``zlib`` / ``bzp2`` blocks on first use; :meth:`BlockArray.read_into` streams the bytes of a block straight
  into a caller buffer (the pinned staging buffers of `gen_cal_image.Pipeline`) without an intermediate array;
* every other tagged node is kept as a :class:`Tagged` mapping / sequence / scalar, so metadata written by
  `roman_datamodels` passes through a read -> write cycle with its tags intact.

It is NOT a validating implementation (no schemas, no extensions, no `lz4`, no streamed or external blocks); where
the real ``asdf`` package is installed, `caltree.open_tree` prefers it.
"""

from __future__ import annotations

import bz2
import hashlib
import io
import os
import re
import struct
import zlib

import numpy as np
import yaml

BLOCK_MAGIC = b"\xd3BLK"
INDEX_HEADER = b"#ASDF BLOCK INDEX"
_ASDF_TAG_PREFIX = "tag:stsci.edu:asdf/"
_NDARRAY_RE = re.compile(r"(^|/)core/ndarray-\d+\.\d+\.\d+$")

_DT_TO_ASDF = {
    "int8": "int8", "int16": "int16", "int32": "int32", "int64": "int64",
    "uint8": "uint8", "uint16": "uint16", "uint32": "uint32", "uint64": "uint64",
    "float16": "float16", "float32": "float32", "float64": "float64",
    "complex64": "complex64", "complex128": "complex128", "bool": "bool8",
}  # fmt: skip
_ASDF_TO_DT = {v: k for k, v in _DT_TO_ASDF.items()}


class AsdfLiteError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------------
# tree node types
# ---------------------------------------------------------------------------------------------------------------
class Tagged:
    """Mixin: a YAML node whose (unknown) tag is carried along."""

    tag: str | None = None


class TaggedDict(dict, Tagged):
    def __init__(self, *a, tag=None, **k):
        super().__init__(*a, **k)
        self.tag = tag


class TaggedList(list, Tagged):
    def __init__(self, *a, tag=None):
        super().__init__(*a)
        self.tag = tag


class TaggedScalar(str, Tagged):
    def __new__(cls, value, tag=None):
        o = super().__new__(cls, value)
        o.tag = tag
        return o


class _Block:
    """One binary block of an open file."""

    __slots__ = ("offset", "data_offset", "allocated", "used", "size", "compression", "checksum", "_cache")

    def __init__(self, offset, data_offset, allocated, used, size, compression, checksum):
        self.offset, self.data_offset = offset, data_offset
        self.allocated, self.used, self.size = allocated, used, size
        self.compression, self.checksum = compression, checksum
        self._cache = None


class BlockArray:
    """Lazy ndarray leaf backed by a block of an open :class:`AsdfLiteFile` (read side)."""

    def __init__(self, owner, source, dtype, shape, offset=0, strides=None):
        self._owner, self.source = owner, int(source)
        self.dtype, self.shape = np.dtype(dtype), tuple(int(s) for s in shape)
        self.offset, self.strides = int(offset), (tuple(strides) if strides is not None else None)

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape, dtype=np.int64)))
    nbytes = property(lambda self: self.size * self.dtype.itemsize)

    def _array(self):
        return self._owner._block_array(self)

    def __array__(self, dtype=None, copy=None):
        a = self._array()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return np.array(a) if copy else a

    def __getitem__(self, key):
        return self._array()[key]

    def __len__(self):
        return self.shape[0]

    def astype(self, dtype, **k):
        return self._array().astype(dtype, **k)

    def is_plain(self):
        """C-contiguous, uncompressed, native little-endian: the block bytes ARE the array."""
        blk = self._owner._blocks[self.source]
        return self.strides is None and not blk.compression and self.dtype.byteorder in ("<", "=", "|")

    def read_into(self, dst, first=0, count=None):
        """Copy ``count`` leading-axis items starting at ``first`` into ``dst`` (any writable C-contiguous buffer, e.g. a
        pinned staging array) straight from the file: one ``readinto`` per call for plain blocks, no temporary."""
        n0 = self.shape[0] if self.shape else 1
        count = n0 - first if count is None else count
        if first < 0 or count < 0 or first + count > n0:
            raise AsdfLiteError("read_into: range outside the leading axis")
        item = (self.nbytes // n0) if n0 else 0
        mv = memoryview(dst).cast("B")
        if mv.nbytes < item * count:
            raise AsdfLiteError(f"read_into: destination holds {mv.nbytes} bytes, need {item * count}")
        if self.is_plain():
            blk = self._owner._blocks[self.source]
            got = os.preadv(self._owner._fd, [mv[: item * count]], blk.data_offset + self.offset + item * first)
            if got != item * count:
                raise AsdfLiteError("read_into: short read")
        else:
            part = self._array()[first : first + count]
            src = np.ascontiguousarray(part, dtype=part.dtype.newbyteorder("="))  # (big-endian blocks: swapped here)
            mv[: item * count] = memoryview(src).cast("B")
        return item * count

    def __repr__(self):
        return f"BlockArray(source={self.source}, dtype={self.dtype}, shape={self.shape})"


# ---------------------------------------------------------------------------------------------------------------
# YAML <-> tree
# ---------------------------------------------------------------------------------------------------------------
def _make_loader(owner):
    class Loader(yaml.SafeLoader):
        pass

    def ndarray(loader, node):
        d = loader.construct_mapping(node, deep=True)
        bo = d.get("byteorder", "little")
        dt = d.get("datatype")
        if not isinstance(dt, str) or dt not in _ASDF_TO_DT:
            raise AsdfLiteError(f"unsupported ndarray datatype {dt!r}")
        npdt = np.dtype(_ASDF_TO_DT[dt])
        if npdt.itemsize > 1:
            npdt = npdt.newbyteorder("<" if bo == "little" else ">")
        if "data" in d:  # inline array
            return np.array(d["data"], dtype=npdt)
        src = d.get("source")
        if not isinstance(src, int):
            raise AsdfLiteError("external / streamed ndarray sources are not supported")
        return BlockArray(owner, src, npdt, d.get("shape", ()), d.get("offset", 0), d.get("strides"))

    def generic(loader, suffix, node):
        tag = node.tag
        if _NDARRAY_RE.search(tag):
            return ndarray(loader, node)
        if isinstance(node, yaml.MappingNode):
            out = TaggedDict(tag=tag)
            out.update(loader.construct_mapping(node, deep=True))
            return out
        if isinstance(node, yaml.SequenceNode):
            return TaggedList(loader.construct_sequence(node, deep=True), tag=tag)
        return TaggedScalar(loader.construct_scalar(node), tag=tag)

    Loader.add_multi_constructor("tag:stsci.edu:asdf/", generic)
    Loader.add_multi_constructor("asdf://", generic)
    Loader.add_multi_constructor("tag:", generic)
    Loader.add_multi_constructor("!", generic)
    return Loader


class _Dumper(yaml.SafeDumper):
    def ignore_aliases(self, data):  # never emit anchors: arrays referenced twice are written twice
        return True


def _short_tag(tag):
    return tag


def _rep_tdict(d, data):
    return d.represent_mapping(data.tag or "tag:yaml.org,2002:map", dict(data))


def _rep_tlist(d, data):
    return d.represent_sequence(data.tag or "tag:yaml.org,2002:seq", list(data))


def _rep_tscalar(d, data):
    return d.represent_scalar(data.tag or "tag:yaml.org,2002:str", str(data))


_Dumper.add_representer(TaggedDict, _rep_tdict)
_Dumper.add_representer(TaggedList, _rep_tlist)
_Dumper.add_representer(TaggedScalar, _rep_tscalar)
_Dumper.add_multi_representer(np.integer, lambda d, v: d.represent_int(int(v)))
_Dumper.add_multi_representer(np.floating, lambda d, v: d.represent_float(float(v)))
_Dumper.add_representer(np.bool_, lambda d, v: d.represent_bool(bool(v)))
_Dumper.add_representer(tuple, lambda d, v: d.represent_sequence("tag:yaml.org,2002:seq", list(v)))


# ---------------------------------------------------------------------------------------------------------------
# reading
# ---------------------------------------------------------------------------------------------------------------
class AsdfLiteFile:
    """``with open_file(path) as f: f["roman"]["data"]`` - the subset of ``asdf.AsdfFile`` the hot path uses."""

    def __init__(self, path):
        self.path = os.fspath(path)
        self._fd = os.open(self.path, os.O_RDONLY)
        self._maps = {}
        try:
            self._parse()
        except Exception:
            self.close()
            raise

    # -- structure --------------------------------------------------------------------------------------------
    def _parse(self):
        size = os.fstat(self._fd).st_size
        head = os.pread(self._fd, min(size, 1 << 16), 0)
        if not head.startswith(b"#ASDF"):
            raise AsdfLiteError(f"{self.path}: not an ASDF file")
        # the YAML document ends with a line holding "..." ; read more if the tree is larger than the first chunk
        buf, pos = head, 0
        m = None
        while True:
            m = re.search(rb"\r?\n\.\.\.\r?\n", buf)
            if m or len(buf) >= size:
                break
            buf += os.pread(self._fd, min(1 << 22, size - len(buf)), len(buf))
        if m:
            yaml_bytes, pos = buf[: m.end()], m.end()
        elif BLOCK_MAGIC in buf:  # (a file without a tree is legal)
            pos = buf.index(BLOCK_MAGIC)
            yaml_bytes = buf[:pos]
        else:
            yaml_bytes, pos = buf, len(buf)
        self._blocks = []
        self._scan_blocks(pos, size)
        text = yaml_bytes.decode("utf-8")
        doc = text[text.index("---") :] if "---" in text else ""
        pre = text[: text.index("---")] if "---" in text else text
        directives = "".join(ln + "\n" for ln in pre.splitlines() if ln.startswith("%"))
        self.tree = yaml.load(directives + doc, Loader=_make_loader(self)) if doc else {}  # noqa: S506 (SafeLoader subclass)
        if self.tree is None:
            self.tree = {}

    def _scan_blocks(self, pos, size):
        # padding may sit between the tree and the first block; blocks are then contiguous via `allocated`
        probe = os.pread(self._fd, min(1 << 16, size - pos), pos)
        k = probe.find(BLOCK_MAGIC)
        if k < 0:
            return
        pos += k
        while pos + 6 <= size:
            hd = os.pread(self._fd, 6, pos)
            if hd[:4] != BLOCK_MAGIC:
                break
            (hsize,) = struct.unpack(">H", hd[4:6])
            if hsize < 48:
                raise AsdfLiteError(f"{self.path}: block header of {hsize} bytes")
            h = os.pread(self._fd, hsize, pos + 6)
            flags, comp, alloc, used, dsize = struct.unpack(">I4sQQQ", h[:32])
            if flags & 1:
                raise AsdfLiteError("streamed blocks are not supported")
            comp = comp.rstrip(b"\0").decode("ascii")
            self._blocks.append(_Block(pos, pos + 6 + hsize, alloc, used, dsize, comp, h[32:48]))
            pos += 6 + hsize + alloc

    # -- payload ----------------------------------------------------------------------------------------------
    def _block_bytes(self, blk):
        """decoded payload of a block as a 1-D uint8 array (memory map for uncompressed blocks)."""
        if blk._cache is not None:
            return blk._cache
        if not blk.compression:
            if blk.used == 0:
                arr = np.zeros(0, np.uint8)
            else:
                arr = np.memmap(self.path, dtype=np.uint8, mode="r", offset=blk.data_offset, shape=(blk.used,))
        else:
            raw = os.pread(self._fd, blk.used, blk.data_offset)
            if blk.compression == "zlib":
                dec = zlib.decompress(raw)
            elif blk.compression == "bzp2":
                dec = bz2.decompress(raw)
            else:
                raise AsdfLiteError(f"block compression {blk.compression!r} is not supported (re-save the file uncompressed)")
            if len(dec) != blk.size:
                raise AsdfLiteError("decompressed block has the wrong size")
            arr = np.frombuffer(dec, dtype=np.uint8)
        blk._cache = arr
        return arr

    def _block_array(self, ref):
        if ref.source < 0 or ref.source >= len(self._blocks):
            raise AsdfLiteError(f"ndarray refers to block {ref.source}, file has {len(self._blocks)}")
        raw = self._block_bytes(self._blocks[ref.source])
        if ref.strides is not None:
            flat = np.frombuffer(raw, dtype=ref.dtype, offset=ref.offset, count=(len(raw) - ref.offset) // ref.dtype.itemsize)
            return np.lib.stride_tricks.as_strided(flat, shape=ref.shape, strides=ref.strides, writeable=False)
        n = ref.size
        if ref.offset + n * ref.dtype.itemsize > len(raw):
            raise AsdfLiteError("ndarray extends beyond its block")
        a = raw[ref.offset : ref.offset + n * ref.dtype.itemsize].view(ref.dtype).reshape(ref.shape)
        return a

    def verify_checksums(self):
        """MD5 of every block against its header (all-zero checksums are 'unchecked' per the standard)."""
        for i, blk in enumerate(self._blocks):
            if blk.checksum == b"\0" * 16:
                continue
            raw = os.pread(self._fd, blk.used, blk.data_offset)
            if hashlib.md5(raw).digest() != blk.checksum:  # noqa: S324 (the format's checksum, not security)
                raise AsdfLiteError(f"{self.path}: checksum mismatch in block {i}")

    # -- mapping interface ------------------------------------------------------------------------------------
    def __getitem__(self, key):
        return self.tree[key]

    def __contains__(self, key):
        return key in self.tree

    def keys(self):
        return self.tree.keys()

    def info(self, max_rows=None):  # (oututils.py:44 prints this)
        return f"<asdf_lite {self.path}: {len(self._blocks)} blocks, top-level keys {list(self.tree)}>"

    def close(self):
        for blk in getattr(self, "_blocks", []):
            blk._cache = None
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def open_file(path):
    return AsdfLiteFile(path)


# ---------------------------------------------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------------------------------------------
def _collect(node, blocks):
    """Replace ndarray leaves by ``!core/ndarray`` mappings that refer to blocks (in document order)."""
    if isinstance(node, BlockArray):
        node = np.asarray(node)
    if isinstance(node, np.ndarray):
        if node.dtype.name not in _DT_TO_ASDF:
            raise AsdfLiteError(f"cannot write arrays of dtype {node.dtype}")
        a = np.ascontiguousarray(node)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        blocks.append(a)
        return TaggedDict(
            {"source": len(blocks) - 1, "datatype": _DT_TO_ASDF[a.dtype.name], "byteorder": "little", "shape": list(a.shape)},
            tag=_ASDF_TAG_PREFIX + "core/ndarray-1.0.0")  # fmt: skip
    if isinstance(node, dict):
        out = TaggedDict(tag=getattr(node, "tag", None))
        for k, v in node.items():
            out[k] = _collect(v, blocks)
        return out
    if isinstance(node, (list, tuple)):
        return TaggedList([_collect(v, blocks) for v in node], tag=getattr(node, "tag", None))
    return node


def write_file(path, tree, checksum=True):
    """Write ``tree`` (dicts / lists / scalars / NumPy arrays / nodes read by this module) as an ASDF 1.x file with
    uncompressed internal blocks and a block index.  Returns the number of bytes written."""
    blocks = []
    body = _collect(tree, blocks)
    top = TaggedDict(body if isinstance(body, dict) else {"data": body}, tag=_ASDF_TAG_PREFIX + "core/asdf-1.1.0")
    out = io.BytesIO()
    out.write(b"#ASDF 1.0.0\n#ASDF_STANDARD 1.5.0\n%YAML 1.1\n%TAG ! tag:stsci.edu:asdf/\n")
    text = yaml.dump(top, Dumper=_Dumper, default_flow_style=None, explicit_start=True, explicit_end=True,
                     allow_unicode=True, width=120, tags={"!": _ASDF_TAG_PREFIX})  # fmt: skip
    # (PyYAML repeats the directives it was given; keep only the document)
    text = text[text.index("---") :]
    out.write(text.encode("utf-8"))
    head = out.getvalue()
    offsets = []
    pos = len(head)
    with open(path, "wb") as f:
        f.write(head)
        for a in blocks:
            raw = memoryview(a).cast("B") if a.size else memoryview(b"")
            md5 = hashlib.md5(raw).digest() if checksum else b"\0" * 16  # noqa: S324
            hdr = struct.pack(">I4sQQQ", 0, b"\0\0\0\0", a.nbytes, a.nbytes, a.nbytes) + md5
            offsets.append(pos)
            f.write(BLOCK_MAGIC + struct.pack(">H", len(hdr)) + hdr)
            f.write(raw)
            pos += 6 + len(hdr) + a.nbytes
        if blocks:
            idx = INDEX_HEADER + b"\n%YAML 1.1\n---\n" + b"".join(b"- %d\n" % o for o in offsets) + b"...\n"
            f.write(idx)
            pos += len(idx)
    return pos
