"""Native reader / writer of the OpenVDB ``.vdb`` file format for the grids PlenVDB stores:
FloatGrid (``Tree_float_5_4_3``) and Vec3SGrid (``Tree_vec3s_5_4_3``), several grids per file.

The reference calls libopenvdb (`openvdb::io::File::write / readGrid`, plenvdb/lib/vdb/plenvdb.h:126-148, 211-240),
which cannot be built here.  This module restates the on-disk layout of OpenVDB 9.1 (file format version 224) from its
sources; every function cites the code it follows (paths relative to openvdb/openvdb/openvdb/ in the reference tree):

  io/Archive.cc:1020-1060  header: magic 0x56444220 (int64), file version, library major/minor, has-grid-offsets, 36-char uuid
  MetaMap.cc / Metadata.h  metadata: count, then (name, type name, uint32 size, payload)
  io/Archive.cc:1248-1430  grid count; per grid: descriptor (unique name, grid type, instance parent; io/GridDescriptor.cc:54-71),
                           three int64 stream offsets (grid, blocks, end), uint32 compression flags, grid metadata, transform,
                           topology, buffers
  math/Transform.cc:151-159, math/Maps.h:836-843  transform: map type name + map payload (UniformScaleMap: five Vec3d)
  tree/Tree.h:1133-1137, tree/RootNode.h:2288-2319, tree/InternalNode.h:2213-2233, tree/LeafNode.h (writeTopology)
                           topology: buffer count 1; root: background, #tiles, #children, tiles, then children in Coord
                           order each as origin + node; internal node: child mask, value mask, compressed tile values, then
                           its children in offset order; leaf: value mask
  tree/LeafNode.h:1438-1447  buffers: leaves in the same depth-first order: value mask again + compressed values
  io/Compression.h:80-160, 645-749  active-mask compression (one metadata byte, up to two inactive values, selection mask,
                           only the active values stored) and io/Compression.cc:79-110 the zlib chunk format
                           (int64 size; <= 0 means that many raw bytes follow)

  io/Compression.cc:172-187, 266-307  the Blosc chunk format (int64 size; <= 0 means raw bytes): one Blosc-1 chunk per buffer,
                           written with blosc_compress_ctx(clevel 9, byte shuffle, typesize 4, LZ4, blocksize = buffer size)

Supported on read: compression NONE / ZIP / BLOSC with or without ACTIVE_MASK, half-float grids, tiles at any level (expanded
into voxels), file versions >= 222.  c-blosc is not vendored in the reference tree and not installed here, so the Blosc-1
chunk container (16-byte header, block offsets, per-block byte-shuffle splits) and the LZ4 block format are restated from
their published specifications (c-blosc README_HEADER.rst / blosc.c: blosc_d; lz4_Block_format.md); chunks compressed
with LZ4 / LZ4HC (what OpenVDB writes) or zlib, stored raw ("memcpyed") or byte-shuffled are decoded, BloscLZ / Snappy / Zstd
and bit-shuffled chunks are rejected with a clear error.
Written: version 224, ZIP | ACTIVE_MASK by default (what OpenVDB itself writes when built without Blosc).

PARITY UNPINNED: there is no OpenVDB build and no ``.vdb`` sample in this environment, so the codec is checked by
round trips and by byte-level structure tests against the layout above, not against files produced by libopenvdb.
"""
import struct
import uuid
import zlib

import numpy as np

MAGIC = 0x56444220
FILE_VERSION = 224
LIB_MAJOR, LIB_MINOR = 9, 1
COMPRESS_NONE, COMPRESS_ZIP, COMPRESS_ACTIVE_MASK, COMPRESS_BLOSC = 0, 1, 2, 4
# io/Compression.h:68-76
NO_MASK_OR_INACTIVE_VALS, NO_MASK_AND_MINUS_BG, NO_MASK_AND_ONE_INACTIVE_VAL, MASK_AND_NO_INACTIVE_VALS, \
    MASK_AND_ONE_INACTIVE_VAL, MASK_AND_TWO_INACTIVE_VALS, NO_MASK_AND_ALL_VALS = range(7)

GRID_TYPES = {"Tree_float_5_4_3": 1, "Tree_vec3s_5_4_3": 3}
_LOG2 = (5, 4, 3)            # upper, lower, leaf
_TOTAL = (12, 7, 3)          # log2 of the voxel span of a node at each level


class VdbError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ Blosc-1 chunks (read only)
def lz4_block_decode(src, out_size):
    """One LZ4 block (lz4_Block_format.md): sequences of token (literal length << 4 | match length - 4), optional 255-run
    length extensions, literals, 2-byte little-endian offset; the last sequence ends after its literals."""
    src = bytes(src)
    n = len(src)
    dst = bytearray(out_size)
    i = o = 0
    while i < n:
        tok = src[i]
        i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                if i >= n:
                    raise VdbError("LZ4 block: truncated literal length")
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        if i + lit > n or o + lit > out_size:
            raise VdbError("LZ4 block: literals run past the end of the buffer")
        dst[o:o + lit] = src[i:i + lit]
        i += lit
        o += lit
        if i >= n:
            break
        if i + 2 > n:
            raise VdbError("LZ4 block: truncated match offset")
        off = src[i] | (src[i + 1] << 8)
        i += 2
        ml = tok & 15
        if ml == 15:
            while True:
                if i >= n:
                    raise VdbError("LZ4 block: truncated match length")
                b = src[i]
                i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        m = o - off
        if off == 0 or m < 0 or o + ml > out_size:
            raise VdbError("LZ4 block: bad match (offset %d, length %d at %d)" % (off, ml, o))
        if off >= ml:
            dst[o:o + ml] = dst[m:m + ml]
        else:                                   # overlapping match: the last `off` bytes repeat
            pat = bytes(dst[m:o])
            dst[o:o + ml] = (pat * (ml // off + 1))[:ml]
        o += ml
    if o != out_size:
        raise VdbError("LZ4 block: decoded %d bytes, expected %d" % (o, out_size))
    return bytes(dst)


_BLOSC_SHUFFLE, _BLOSC_MEMCPYED, _BLOSC_BITSHUFFLE, _BLOSC_DONT_SPLIT = 0x1, 0x2, 0x4, 0x10
_BLOSC_MAX_SPLITS, _BLOSC_MIN_BUFFERSIZE = 16, 128
_BLOSC_FORMATS = {0: "blosclz", 1: "lz4", 2: "snappy", 3: "zlib", 4: "zstd"}


def blosc_decompress(buf, expect_nbytes=None):
    """One Blosc-1 chunk -> bytes.  Header (c-blosc README_HEADER.rst): version, versionlz, flags, typesize, then uint32
    nbytes, blocksize, cbytes; then one int32 offset per block; a block holds `typesize` byte-shuffle splits (or one), each an
    int32 compressed size followed by that many bytes — stored raw when the size equals the split's length (blosc.c: blosc_d)."""
    buf = bytes(buf)
    if len(buf) < 16:
        raise VdbError("Blosc chunk: shorter than its 16-byte header")
    version, _versionlz, flags, typesize = buf[0], buf[1], buf[2], buf[3]
    nbytes, blocksize, cbytes = struct.unpack_from("<III", buf, 4)
    if version != 2:
        raise VdbError("Blosc chunk: format version %d is not supported (expected 2)" % version)
    if expect_nbytes is not None and nbytes != expect_nbytes:
        raise VdbError("Blosc chunk: holds %d bytes, expected %d" % (nbytes, expect_nbytes))
    if cbytes > len(buf):
        raise VdbError("Blosc chunk: header says %d compressed bytes, only %d present" % (cbytes, len(buf)))
    if nbytes == 0:
        return b""
    if flags & _BLOSC_MEMCPYED:
        if 16 + nbytes > len(buf):
            raise VdbError("Blosc chunk: truncated raw payload")
        return buf[16:16 + nbytes]
    if flags & _BLOSC_BITSHUFFLE:
        raise VdbError("Blosc chunk: bit-shuffled chunks are not supported")
    fmt = flags >> 5
    if fmt not in (1, 3):
        raise VdbError("Blosc chunk: compressor '%s' is not supported (LZ4 and zlib are)" % _BLOSC_FORMATS.get(fmt, fmt))
    if blocksize == 0 or typesize == 0:
        raise VdbError("Blosc chunk: zero block size / type size")
    nblocks = (nbytes + blocksize - 1) // blocksize
    if 16 + 4 * nblocks > len(buf):
        raise VdbError("Blosc chunk: truncated block offsets")
    bstarts = struct.unpack_from("<%di" % nblocks, buf, 16)
    out = bytearray(nbytes)
    for j in range(nblocks):
        bsize, leftover = blocksize, False
        if j == nblocks - 1 and nbytes % blocksize:
            bsize, leftover = nbytes % blocksize, True
        split = (not (flags & _BLOSC_DONT_SPLIT) and typesize <= _BLOSC_MAX_SPLITS and bsize // typesize >= _BLOSC_MIN_BUFFERSIZE
                 and not leftover)
        nsplits = typesize if split else 1
        neblock = bsize // nsplits
        p = bstarts[j]
        parts = []
        for _ in range(nsplits):
            if p < 16 or p + 4 > len(buf):
                raise VdbError("Blosc chunk: block offset out of range")
            c = struct.unpack_from("<i", buf, p)[0]
            p += 4
            if c < 0 or p + c > len(buf):
                raise VdbError("Blosc chunk: split of %d bytes runs past the end" % c)
            if c == neblock:
                parts.append(buf[p:p + c])
            elif fmt == 1:
                parts.append(lz4_block_decode(buf[p:p + c], neblock))
            else:
                d = zlib.decompress(buf[p:p + c])
                if len(d) != neblock:
                    raise VdbError("Blosc chunk: zlib split decoded to %d bytes, expected %d" % (len(d), neblock))
                parts.append(d)
            p += c
        block = b"".join(parts)
        if len(block) != bsize:
            raise VdbError("Blosc chunk: block of %d bytes, expected %d" % (len(block), bsize))
        if (flags & _BLOSC_SHUFFLE) and typesize > 1:
            ne = bsize // typesize                      # byte j of element i sits at j * ne + i; the tail bytes are not shuffled
            body = np.frombuffer(block, np.uint8, ne * typesize).reshape(typesize, ne).T.tobytes()
            block = body + block[ne * typesize:]
        out[j * blocksize:j * blocksize + bsize] = block
    return bytes(out)


# ------------------------------------------------------------------------------------------------ byte stream helpers
class _Writer:
    def __init__(self):
        self.parts, self.n = [], 0

    def raw(self, b):
        self.parts.append(bytes(b))
        self.n += len(b)

    def pack(self, fmt, *v):
        self.raw(struct.pack("<" + fmt, *v))

    def string(self, s):                 # io/io.h writeString: uint32 length + characters
        b = s.encode("utf-8")
        self.pack("I", len(b))
        self.raw(b)

    def tell(self):
        return self.n

    def getvalue(self):
        return b"".join(self.parts)


class _Reader:
    def __init__(self, data):
        self.d, self.p = memoryview(data), 0

    def raw(self, n):
        if self.p + n > len(self.d):
            raise VdbError("truncated .vdb file")
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def unpack(self, fmt):
        s = struct.Struct("<" + fmt)
        v = s.unpack(self.raw(s.size))
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.unpack("I")
        return bytes(self.raw(n)).decode("utf-8", "replace")


# ------------------------------------------------------------------------------------------------ metadata
def _meta_bytes(kind, value):
    if kind == "string":
        return value.encode("utf-8")
    if kind == "bool":
        return struct.pack("<?", bool(value))
    if kind == "int32":
        return struct.pack("<i", int(value))
    if kind == "int64":
        return struct.pack("<q", int(value))
    if kind == "float":
        return struct.pack("<f", float(value))
    if kind == "vec3i":
        return struct.pack("<3i", *[int(v) for v in value])
    raise VdbError("unsupported metadata type " + kind)


def _write_metamap(w, metas):
    """MetaMap::writeMeta: Index32 count; the map is ordered by name (std::map)."""
    items = sorted(metas.items())
    w.pack("I", len(items))
    for name, (kind, value) in items:
        w.string(name)
        w.string(kind)
        payload = _meta_bytes(kind, value)
        w.pack("I", len(payload))          # Metadata::write: size, then value
        w.raw(payload)


def _read_metamap(r):
    out = {}
    for _ in range(r.unpack("I")):
        name, kind = r.string(), r.string()
        payload = bytes(r.raw(r.unpack("I")))
        if kind == "string":
            out[name] = payload.decode("utf-8", "replace")
        elif kind == "bool":
            out[name] = bool(payload[0]) if payload else False
        elif kind == "int32":
            out[name] = struct.unpack("<i", payload)[0]
        elif kind == "int64":
            out[name] = struct.unpack("<q", payload)[0]
        elif kind == "float":
            out[name] = struct.unpack("<f", payload)[0]
        elif kind == "double":
            out[name] = struct.unpack("<d", payload)[0]
        elif kind == "vec3i":
            out[name] = struct.unpack("<3i", payload)
        else:
            out[name] = payload                 # unknown types (e.g. __delayedload) are carried as bytes
    return out


# ------------------------------------------------------------------------------------------------ transforms
_MAP_BYTES = {"UniformScaleMap": 120, "ScaleMap": 120, "TranslationMap": 24, "UniformScaleTranslateMap": 144, "ScaleTranslateMap": 144,
              "AffineMap": 128, "UnitaryMap": 128}


def _write_identity_transform(w):
    """Grid's default transform: createLinearTransform(1.0) = UniformScaleMap(1) (math/Maps.h:836-843: scale, voxel size,
    inverse scale, inverse scale squared, inverse twice scale)."""
    w.string("UniformScaleMap")
    for v in (1.0, 1.0, 1.0, 1.0, 0.5):
        w.pack("3d", v, v, v)


def _read_transform(r):
    kind = r.string()
    if kind == "NonlinearFrustumMap":         # bbox, taper, depth, second map (math/Maps.h:2424-2432)
        r.raw(48 + 16)
        inner = r.string()
        r.raw(_MAP_BYTES.get(inner, 128))
        return kind, None
    if kind not in _MAP_BYTES:
        raise VdbError("unknown transform map type %r" % kind)
    payload = bytes(r.raw(_MAP_BYTES[kind]))
    scale = struct.unpack("<3d", payload[24:48] if kind.endswith("TranslateMap") else payload[:24]) if "Scale" in kind else (1.0, 1.0, 1.0)
    return kind, scale


# ------------------------------------------------------------------------------------------------ value chunks
def _write_data(w, arr, compression):
    """io/Compression.h:363-372 writeData + io/Compression.cc:79-110 zipToStream."""
    b = np.ascontiguousarray(arr).tobytes()
    if compression & COMPRESS_BLOSC:
        raise VdbError("Blosc output is not supported")
    if compression & COMPRESS_ZIP:
        z = zlib.compress(b, zlib.Z_DEFAULT_COMPRESSION) if b else b""
        if b and len(z) < len(b):
            w.pack("q", len(z))
            w.raw(z)
        else:
            w.pack("q", -len(b))
            w.raw(b)
    else:
        w.raw(b)


def _read_data(r, dtype, count, comps, compression):
    nbytes = int(count) * comps * np.dtype(dtype).itemsize
    if compression & COMPRESS_BLOSC:
        n = r.unpack("q")
        if n <= 0:                                  # bloscToStream stores incompressible data raw, like zipToStream
            if -n != nbytes:
                raise VdbError("expected a %d-byte chunk, got %d" % (nbytes, -n))
            return np.frombuffer(bytes(r.raw(-n)), dtype=dtype).reshape(count, comps).copy()
        b = blosc_decompress(r.raw(n), nbytes)      # io/Compression.cc:285-304
        return np.frombuffer(b, dtype=dtype).reshape(count, comps).copy()
    if compression & COMPRESS_ZIP:
        n = r.unpack("q")
        if n <= 0:
            if -n != nbytes:
                raise VdbError("expected a %d-byte chunk, got %d" % (nbytes, -n))
            b = bytes(r.raw(-n))
        else:
            b = zlib.decompress(bytes(r.raw(n)))
            if len(b) != nbytes:
                raise VdbError("expected to decompress %d bytes, got %d" % (nbytes, len(b)))
    else:
        b = bytes(r.raw(nbytes))
    return np.frombuffer(b, dtype=dtype).reshape(count, comps).copy()


def _mask_compress(vals, vmask, cmask, background):
    """io/Compression.h:87-160 MaskCompress: scan the inactive, non-child values in order for up to three distinct ones."""
    inactive = [background.copy(), background.copy()]
    n_unique = 0
    idx = np.nonzero(~vmask & ~cmask)[0]
    if idx.size:
        # distinct values in order of first appearance, capped at three
        rows = vals[idx]
        _, first = np.unique(rows, axis=0, return_index=True)
        for i in np.sort(first)[:3]:
            if n_unique < 2:
                inactive[n_unique] = rows[i].copy()
            n_unique += 1
    eq = lambda a, b: bool(np.array_equal(a, b))
    meta = NO_MASK_OR_INACTIVE_VALS
    if n_unique == 1:
        if not eq(inactive[0], background):
            meta = NO_MASK_AND_MINUS_BG if eq(inactive[0], -background) else NO_MASK_AND_ONE_INACTIVE_VAL
    elif n_unique == 2:
        if not eq(inactive[0], background) and not eq(inactive[1], background):
            meta = MASK_AND_TWO_INACTIVE_VALS
        elif eq(inactive[1], background):
            meta = MASK_AND_NO_INACTIVE_VALS if eq(inactive[0], -background) else MASK_AND_ONE_INACTIVE_VAL
        elif eq(inactive[0], background):
            meta = MASK_AND_NO_INACTIVE_VALS if eq(inactive[1], -background) else MASK_AND_ONE_INACTIVE_VAL
            inactive[0], inactive[1] = inactive[1], inactive[0]
    elif n_unique > 2:
        meta = NO_MASK_AND_ALL_VALS
    return meta, inactive


def _pack_mask(bits):
    return np.packbits(bits.astype(np.uint8), bitorder="little").tobytes()       # NodeMask::save: the 64-bit words, bit n = word n>>6


def _unpack_mask(b, n):
    return np.unpackbits(np.frombuffer(b, np.uint8), bitorder="little")[:n].astype(bool)


def _write_values(w, vals, vmask, cmask, background, compression):
    """io/Compression.h:645-749 writeCompressedValues.  vals [n, comps] float32."""
    if not compression & COMPRESS_ACTIVE_MASK:
        w.pack("b", NO_MASK_AND_ALL_VALS)
        _write_data(w, vals, compression)
        return
    meta, inactive = _mask_compress(vals, vmask, cmask, background)
    w.pack("b", meta)
    if meta in (NO_MASK_AND_ONE_INACTIVE_VAL, MASK_AND_ONE_INACTIVE_VAL, MASK_AND_TWO_INACTIVE_VALS):
        w.raw(inactive[0].astype(np.float32).tobytes())
        if meta == MASK_AND_TWO_INACTIVE_VALS:
            w.raw(inactive[1].astype(np.float32).tobytes())
    if meta == NO_MASK_AND_ALL_VALS:
        _write_data(w, vals, compression)
        return
    if meta in (MASK_AND_NO_INACTIVE_VALS, MASK_AND_ONE_INACTIVE_VAL, MASK_AND_TWO_INACTIVE_VALS):
        sel = ~vmask & (vals == inactive[1]).all(1)
        w.raw(_pack_mask(sel))
    _write_data(w, vals[vmask], compression)


def _read_values(r, n, comps, vmask, background, compression, half, version):
    """io/Compression.h readCompressedValues (the mirror image of the above)."""
    dtype = np.float16 if half else np.float32
    meta = r.unpack("b") if version >= 222 else NO_MASK_AND_ALL_VALS
    inactive0 = background if meta == NO_MASK_OR_INACTIVE_VALS else -background
    inactive1 = background
    if meta in (NO_MASK_AND_ONE_INACTIVE_VAL, MASK_AND_ONE_INACTIVE_VAL, MASK_AND_TWO_INACTIVE_VALS):
        inactive0 = np.frombuffer(bytes(r.raw(4 * comps)), np.float32).copy()      # stored as a full ValueT even for half grids
        if meta == MASK_AND_TWO_INACTIVE_VALS:
            inactive1 = np.frombuffer(bytes(r.raw(4 * comps)), np.float32).copy()
    sel = None
    if meta in (MASK_AND_NO_INACTIVE_VALS, MASK_AND_ONE_INACTIVE_VAL, MASK_AND_TWO_INACTIVE_VALS):
        sel = _unpack_mask(bytes(r.raw(n // 8)), n)
    mask_compressed = bool(compression & COMPRESS_ACTIVE_MASK) and meta != NO_MASK_AND_ALL_VALS and version >= 222
    count = int(vmask.sum()) if mask_compressed else n
    data = _read_data(r, dtype, count, comps, compression).astype(np.float32)
    if not mask_compressed:
        return data
    out = np.empty((n, comps), np.float32)
    out[:] = inactive0
    if sel is not None:
        out[sel] = inactive1
    out[vmask] = data
    return out


# ------------------------------------------------------------------------------------------------ writing
def _node_tables(topo):
    """Host tables of a plenvdb_b200 Topology as (upper origins sorted in Coord order, per-upper child table, lower tables)."""
    keys = topo.h_root_keys[: topo.n_upper].astype(np.uint64)
    ox = ((keys >> np.uint64(42)) & np.uint64(0x1FFFFF)).astype(np.int64) << 12
    oy = ((keys >> np.uint64(21)) & np.uint64(0x1FFFFF)).astype(np.int64) << 12
    oz = (keys & np.uint64(0x1FFFFF)).astype(np.int64) << 12
    org = np.stack([ox, oy, oz], 1)
    order = np.lexsort((org[:, 2], org[:, 1], org[:, 0]))       # RootNode's std::map<Coord,...>: x, then y, then z
    return org, order


def encode_grids(topo, planes, compression=COMPRESS_ZIP | COMPRESS_ACTIVE_MASK):
    """-> bytes of a .vdb file.  `planes`: list of (grid name, float32 array [n_leaf, 512, C] with C in {1, 3}); all grids
    share `topo` (a plenvdb_b200 Topology: host child tables, leaf origins, 512-bit leaf masks)."""
    w = _Writer()
    # ---- Archive::writeHeader
    w.pack("q", MAGIC)
    w.pack("I", FILE_VERSION)
    w.pack("II", LIB_MAJOR, LIB_MINOR)
    w.pack("b", 1)                                     # seekable: grid offsets present
    w.raw(str(uuid.uuid4()).encode("ascii"))           # os << boost uuid: 36 ASCII characters
    _write_metamap(w, {})                              # file-level metadata
    w.pack("i", len(planes))
    org, order = _node_tables(topo)
    upper = topo.h_upper.reshape(-1, 32768)
    lower = topo.h_lower.reshape(-1, 4096)
    lmask = np.unpackbits(topo.h_leaf_mask[: max(topo.n_leaf, 1)].view(np.uint8).reshape(-1, 64), axis=1, bitorder="little").astype(bool)
    n_active = int(lmask[: topo.n_leaf].sum())
    if topo.n_leaf:
        lo3 = topo.h_leaf_origin[: topo.n_leaf]
        bmin, bmax = lo3.min(0), lo3.max(0) + 7
    else:
        bmin, bmax = np.zeros(3, np.int64), np.zeros(3, np.int64)
    for name, plane in planes:
        comps = plane.shape[-1]
        if comps not in (1, 3):
            raise VdbError("a grid has 1 (float) or 3 (vec3s) components, got %d" % comps)
        gtype = "Tree_float_5_4_3" if comps == 1 else "Tree_vec3s_5_4_3"
        background = np.zeros(comps, np.float32)
        # ---- GridDescriptor::writeHeader + writeStreamPos (patched below)
        w.string(name)
        w.string(gtype)
        w.string("")                                   # instance parent
        pos_at = w.tell()
        w.pack("3q", 0, 0, 0)
        grid_pos = w.tell()
        w.pack("I", compression)                       # Archive::setGridCompression (grid class unknown: flags unchanged)
        metas = {"class": ("string", "unknown"), "file_compressor": ("string", _compression_name(compression)),
                 "file_bbox_min": ("vec3i", bmin), "file_bbox_max": ("vec3i", bmax),
                 "file_mem_bytes": ("int64", int(plane.size) * 4), "file_voxel_count": ("int64", n_active),
                 "is_local_space": ("bool", False), "is_saved_as_half_float": ("bool", False), "name": ("string", name)}
        if comps == 3:
            metas["vector_type"] = ("string", "invariant")
        _write_metamap(w, metas)
        _write_identity_transform(w)
        # ---- Tree::writeTopology
        w.pack("i", 1)                                 # buffer count
        w.raw(background.tobytes())
        w.pack("II", 0, topo.n_upper)                  # tiles, children
        leaf_order = []
        for u in order:
            w.pack("3i", *[int(v) for v in org[u]])
            cm = upper[u] >= 0
            w.raw(_pack_mask(cm))
            w.raw(_pack_mask(np.zeros(32768, bool)))
            _write_values(w, np.zeros((32768, comps), np.float32), np.zeros(32768, bool), cm, background, compression)
            for lo in upper[u][cm]:
                cml = lower[lo] >= 0
                w.raw(_pack_mask(cml))
                w.raw(_pack_mask(np.zeros(4096, bool)))
                _write_values(w, np.zeros((4096, comps), np.float32), np.zeros(4096, bool), cml, background, compression)
                for lf in lower[lo][cml]:
                    w.raw(_pack_mask(lmask[lf]))       # LeafNode::writeTopology: value mask
                    leaf_order.append(int(lf))
        block_pos = w.tell()
        # ---- Tree::writeBuffers: leaves in the same order
        for lf in leaf_order:
            w.raw(_pack_mask(lmask[lf]))
            _write_values(w, np.ascontiguousarray(plane[lf], np.float32).reshape(512, comps), lmask[lf], np.zeros(512, bool), background, compression)
        end_pos = w.tell()
        w.parts.append(("patch", pos_at, struct.pack("<3q", grid_pos, block_pos, end_pos)))
    # apply the offset patches
    patches = [p for p in w.parts if isinstance(p, tuple)]
    w.parts = [p for p in w.parts if not isinstance(p, tuple)]
    out = bytearray(w.getvalue())
    for _, at, b in patches:
        out[at:at + len(b)] = b
    return bytes(out)


def _compression_name(c):
    names = [n for f, n in ((COMPRESS_ZIP, "zip"), (COMPRESS_BLOSC, "blosc"), (COMPRESS_ACTIVE_MASK, "active values")) if c & f]
    return " + ".join(names) if names else "none"


def write_vdb(path, topo, planes, compression=COMPRESS_ZIP | COMPRESS_ACTIVE_MASK):
    with open(path, "wb") as f:
        f.write(encode_grids(topo, planes, compression))


# ------------------------------------------------------------------------------------------------ reading
def is_vdb(data):
    return len(data) >= 8 and struct.unpack("<q", bytes(data[:8]))[0] == MAGIC


def _expand_tile(origin, log2span, value, coords, values, limit):
    n = 1 << log2span
    if n ** 3 > limit:
        raise VdbError("active tile of %d^3 voxels is too large to expand" % n)
    ax = np.arange(n, dtype=np.int32)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3) + np.asarray(origin, np.int32)
    coords.append(g)
    values.append(np.broadcast_to(value, (g.shape[0], value.shape[0])).copy())


def _read_internal(r, level, origin, comps, background, compression, half, version, leaves, coords, values, tile_limit):
    """InternalNode::readTopology (tree/InternalNode.h:2238-2290) for level 0 (32^3 children of 128^3) / 1 (16^3 of 8^3)."""
    n = 1 << (3 * _LOG2[level])
    cm = _unpack_mask(bytes(r.raw(n // 8)), n)
    vm = _unpack_mask(bytes(r.raw(n // 8)), n)
    vals = _read_values(r, n, comps, vm, background, compression, half, version)
    l2, child_total = _LOG2[level], _TOTAL[level + 1]
    idx = np.arange(n)
    off = np.stack([(idx >> (2 * l2)), (idx >> l2) & ((1 << l2) - 1), idx & ((1 << l2) - 1)], 1).astype(np.int64) << child_total
    for i in np.nonzero(vm & ~cm)[0]:                       # active tiles
        _expand_tile(np.asarray(origin) + off[i], child_total, vals[i], coords, values, tile_limit)
    for i in np.nonzero(cm)[0]:
        o = np.asarray(origin, np.int64) + off[i]
        if level == 0:
            _read_internal(r, 1, o, comps, background, compression, half, version, leaves, coords, values, tile_limit)
        else:
            leaves.append((o.astype(np.int32), _unpack_mask(bytes(r.raw(64)), 512)))     # LeafNode::readTopology: value mask


def decode_grids(data, tile_limit=1 << 24):
    """-> list of dicts {name, type, components, background, coords int32 [n,3], values float32 [n,C], metadata, transform}
    holding the ACTIVE voxels of every grid in the file (tiles expanded)."""
    r = _Reader(data)
    if r.unpack("q") != MAGIC:
        raise VdbError("not a VDB file")
    version = r.unpack("I")
    if version < 222:
        raise VdbError("file format version %d is older than 222 (per-grid compression flags)" % version)
    r.unpack("II")
    has_offsets = r.unpack("b")
    r.raw(36)                                               # uuid
    _read_metamap(r)
    out = []
    for _ in range(r.unpack("i")):
        uname, gtype, parent = r.string(), r.string(), r.string()
        half = gtype.endswith("_HalfFloat")
        if half:
            gtype = gtype[: -len("_HalfFloat")]
        grid_pos, block_pos, end_pos = r.unpack("3q")
        if gtype not in GRID_TYPES:
            if has_offsets and end_pos > 0:
                r.p = end_pos
                continue
            raise VdbError("unsupported grid type %r" % gtype)
        if parent:
            raise VdbError("instanced grids are not supported")
        comps = GRID_TYPES[gtype]
        compression = r.unpack("I")
        meta = _read_metamap(r)
        transform = _read_transform(r)
        if r.unpack("i") != 1:
            raise VdbError("multi-buffer trees are not supported")
        background = np.frombuffer(bytes(r.raw(4 * comps)), np.float32).copy()
        n_tiles, n_children = r.unpack("II")
        coords, values, leaves = [], [], []
        for _ in range(n_tiles):
            o = r.unpack("3i")
            v = np.frombuffer(bytes(r.raw(4 * comps)), np.float32).copy()
            if r.unpack("?"):
                _expand_tile(o, 12, v, coords, values, tile_limit)
        for _ in range(n_children):
            o = r.unpack("3i")
            _read_internal(r, 0, o, comps, background, compression, half, version, leaves, coords, values, tile_limit)
        for o, _topo_mask in leaves:                        # buffers: value mask again, then the values
            vm = _unpack_mask(bytes(r.raw(64)), 512)
            vals = _read_values(r, 512, comps, vm, background, compression, half, version)
            on = np.nonzero(vm)[0]
            if on.size:
                xyz = np.stack([o[0] + (on >> 6), o[1] + ((on >> 3) & 7), o[2] + (on & 7)], 1).astype(np.int32)
                coords.append(xyz)
                values.append(vals[on])
        if has_offsets and end_pos > 0:
            r.p = end_pos
        out.append(dict(name=uname.split("\x1e")[0], type=gtype, components=comps, background=background, metadata=meta, transform=transform,
                        compression=_compression_name(compression),
                        coords=np.concatenate(coords) if coords else np.zeros((0, 3), np.int32),
                        values=np.concatenate(values) if values else np.zeros((0, comps), np.float32)))
    return out


def read_vdb(path):
    with open(path, "rb") as f:
        return decode_grids(f.read())
