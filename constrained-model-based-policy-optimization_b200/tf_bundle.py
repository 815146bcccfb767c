"""Reader (and, for fixtures, writer) of TensorFlow's tensor-bundle files -- the `variables/variables.index` +
`variables/variables.data-00000-of-00001` pair inside the SavedModel that `Logger.save_tf` writes for the policy
(utilities/logx.py:202-259, called by CPOPolicy.save, policies/cpo_policy.py:890-894) -- in plain Python / numpy, so
a policy trained by the reference can drive the B200 rollout without TensorFlow.

PARITY UNPINNED: TensorFlow 1.14 is not installable here, so no file written by the reference exists to test
against.  The format is restated from TensorFlow's published definitions (tensor_bundle.proto, the table format of
tensorflow/core/lib/io/ = LevelDB's table_format.md, crc32c masking of lib/hash/crc32c.h, Snappy's format
description); the tests check the reader against this module's own independent writer, hand-assembled blocks
(prefix compression, restart points, a Snappy-compressed block) and the published CRC-32C check value.

Index file = LevelDB table: data blocks, metaindex block, index block, 48-byte footer (two block handles as
varint64 pairs, padding, magic 0xdb4775248b80fb57).  A block = prefix-compressed entries (shared, unshared,
value_len varint32s + key suffix + value), a uint32 array of restart offsets and its length; followed on disk by
1 byte compression type (0 none, 1 snappy) and the masked CRC-32C of block + type.  Key "" holds BundleHeaderProto
(num_shards, endianness, version); every other key is a tensor name whose value is a BundleEntryProto (dtype,
shape, shard_id, offset, size, masked crc32c of the bytes).
"""
import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- CRC-32C (Castagnoli), table driven ---------------------------------------------------------------------
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()


def crc32c(data):
    tab = _CRC_TABLE
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---- varints / protobuf wire format -------------------------------------------------------------------------
def _get_varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_fields(buf):
    """[(field number, wire type, value)] of one serialized message (value: int or bytes)."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((fno, wt, v))
    return out


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_shape(buf):
    dims = []
    for fno, _, v in _pb_fields(buf):
        if fno == 2:                                       # repeated Dim
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif fno == 3 and v:
            raise ValueError("tensor of unknown rank in a checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for fno, _, v in _pb_fields(buf):
        if fno == 1:
            e["dtype"] = v
        elif fno == 2:
            e["shape"] = _parse_shape(v)
        elif fno == 3:
            e["shard_id"] = v
        elif fno == 4:
            e["offset"] = v
        elif fno == 5:
            e["size"] = v
        elif fno == 6:
            e["crc32c"] = v
        elif fno == 7:
            e["sliced"] = True
    return e


# ---- Snappy (decompression only; the writer's compressor below emits the simplest valid stream) ----------------
def snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):                                 # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ---- table ----------------------------------------------------------------------------------------------------
def _read_block(f, offset, size, verify):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated table block")
    body, ctype = raw[:size], raw[size]
    if verify:
        want = struct.unpack("<I", raw[size + 1:])[0]
        if mask_crc(crc32c(raw[:size + 1])) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        body = snappy_decompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression %d" % ctype)
    return body


def _block_entries(body):
    n_restarts = struct.unpack_from("<I", body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(body, pos)
        unshared, pos = _get_varint(body, pos)
        vlen, pos = _get_varint(body, pos)
        key = key[:shared] + bytes(body[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(body[pos:pos + vlen])))
        pos += vlen
    return out


def _handle(buf, pos):
    off, pos = _get_varint(buf, pos)
    size, pos = _get_varint(buf, pos)
    return off, size, pos


def read_index(index_path, verify=True):
    """{tensor name: entry dict} and the header dict of a bundle's .index file."""
    with open(index_path, "rb") as f:
        f.seek(0, os.SEEK_END)
        total = f.tell()
        if total < 48:
            raise ValueError("%s is too small to be a table" % index_path)
        f.seek(total - 48)
        footer = f.read(48)
        if struct.unpack("<Q", footer[40:])[0] != MAGIC:
            raise ValueError("%s: bad table magic" % index_path)
        _, _, pos = _handle(footer, 0)                      # metaindex (unused)
        ioff, isize, _ = _handle(footer, pos)
        entries = {}
        header = None
        for _, hv in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, bsize, _ = _handle(hv, 0)
            for k, v in _block_entries(_read_block(f, boff, bsize, verify)):
                if k == b"":
                    header = dict(num_shards=1, endianness=0)
                    for fno, _, val in _pb_fields(v):
                        if fno == 1:
                            header["num_shards"] = val
                        elif fno == 2:
                            header["endianness"] = val
                else:
                    entries[k.decode("utf-8")] = _parse_entry(v)
    if header is None:
        raise ValueError("%s has no bundle header" % index_path)
    if header["endianness"] != 0:
        raise ValueError("big-endian bundles are not supported")
    return entries, header


def read_bundle(prefix, names=None, verify=True):
    """{name: numpy array} for the tensors of the bundle `prefix` (.index / .data-xxxxx-of-yyyyy); `names`
    restricts the set (all numeric tensors otherwise; string / variant / sliced entries are skipped)."""
    entries, header = read_index(prefix + ".index", verify)
    out = {}
    files = {}
    try:
        for name, e in entries.items():
            if names is not None and name not in names:
                continue
            if e["dtype"] not in _DTYPES or e["sliced"]:
                if names is not None:
                    raise ValueError("tensor %r has an unsupported dtype / is partitioned" % name)
                continue
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"]), "rb")
            f = files[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            dt = np.dtype(_DTYPES[e["dtype"]])
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if len(raw) != e["size"] or count * dt.itemsize != e["size"]:
                raise ValueError("tensor %r: %d bytes for shape %s %s" % (name, len(raw), e["shape"], dt))
            if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
                raise ValueError("tensor %r: checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for f in files.values():
            f.close()
    if names is not None:
        missing = [n for n in names if n not in out]
        if missing:
            raise KeyError("not in the bundle: %s" % ", ".join(missing))
    return out


# ---- writer (fixtures for the tests; also lets a B200-side policy be handed back in the same container) -------
def _pb_varint_field(fno, v):
    return _put_varint(fno << 3) + _put_varint(v & 0xFFFFFFFFFFFFFFFF)


def _pb_bytes_field(fno, b):
    return _put_varint((fno << 3) | 2) + _put_varint(len(b)) + b


def _snappy_literal_only(data):
    """A valid Snappy stream made of literals (plus one back-reference per run of >= 8 equal bytes, so that the
    reader's copy path is exercised by files this module writes)."""
    out = bytearray(_put_varint(len(data)))
    i, n = 0, len(data)

    def literal(chunk):
        ln = len(chunk) - 1
        if ln < 60:
            out.append(ln << 2)
        else:
            nb = (ln.bit_length() + 7) // 8
            out.append((59 + nb) << 2)
            out.extend(ln.to_bytes(nb, "little"))
        out.extend(chunk)

    start = 0
    while i < n:
        j = i
        while j + 1 < n and data[j + 1] == data[i]:
            j += 1
        run = j - i + 1
        if run >= 8:
            literal(data[start:i + 1])                     # up to and including the first byte of the run
            rem = run - 1
            while rem > 0:
                ln = min(rem, 64)
                out.append(((ln - 1) << 2) | 2)            # copy, 2-byte offset 1 (overlapping)
                out.extend((1).to_bytes(2, "little"))
                rem -= ln
            i = j + 1
            start = i
        else:
            i = j + 1
    if start < n:
        literal(data[start:n])
    return bytes(out)


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_bundle(prefix, tensors, block_size=4096, compress=False):
    """Write {name: numpy array} as a one-shard bundle (`prefix`.index, `prefix`.data-00000-of-00001)."""
    names = sorted(tensors)                                 # table keys must be sorted (bytewise)
    data = bytearray()
    items = [(b"", _pb_varint_field(1, 1) + _pb_bytes_field(3, _pb_varint_field(1, 1)))]   # header: 1 shard, little endian, version
    for name in sorted(names, key=lambda s: s.encode("utf-8")):
        a = np.asarray(tensors[name], order="C")        # (ascontiguousarray would turn a scalar into shape (1,))
        if a.dtype not in _DTYPE_IDS:
            raise ValueError("unsupported dtype %s" % a.dtype)
        raw = a.tobytes()
        shape = b"".join(_pb_bytes_field(2, _pb_varint_field(1, d)) for d in a.shape)
        entry = _pb_varint_field(1, _DTYPE_IDS[a.dtype]) + _pb_bytes_field(2, shape)
        if len(data):
            entry += _pb_varint_field(4, len(data))
        entry += _pb_varint_field(5, len(raw))
        entry += _put_varint((6 << 3) | 5) + struct.pack("<I", mask_crc(crc32c(raw)))
        items.append((name.encode("utf-8"), entry))
        data += raw
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))

    out = bytearray()

    def emit(body):
        ctype = 0
        if compress:
            body, ctype = _snappy_literal_only(body), 1
        off = len(out)
        out.extend(body)
        out.append(ctype)
        out.extend(struct.pack("<I", mask_crc(crc32c(bytes(body) + bytes([ctype])))))
        return _put_varint(off) + _put_varint(len(body))

    index = _BlockBuilder(restart_interval=1)
    blk = _BlockBuilder()
    for k, v in items:
        blk.add(k, v)
        if blk.size() >= block_size:
            index.add(blk.last, emit(blk.finish()))
            blk = _BlockBuilder()
    if len(blk.buf) or not len(index.buf):
        index.add(blk.last, emit(blk.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
