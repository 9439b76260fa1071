"""TensorFlow-free reader / writer for TF "V2" checkpoints (tensor bundles), the format produced by the
reference's `tf.train.Saver().save(...)` and consumed by `saver.restore(sess, checkpoint)`
(train_cnn_networks_hgru.py:188, 248-250, 303-313, 373-383).

A checkpoint `prefix` is two files:

  prefix.index                  an SSTable (LevelDB table format): key "" -> BundleHeaderProto, every other key
                                is a variable name -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
  prefix.data-00000-of-00001    the tensors' raw little-endian bytes at those offsets

Nothing here needs TensorFlow, protobuf or snappy: the few protobuf messages are decoded by hand, and both
uncompressed and snappy-compressed table blocks are understood.  `read_checkpoint(prefix)` returns the flat
`{variable name: ndarray}` dict that `monkey_pose_b200.model.load_params` takes; `write_checkpoint` emits a bundle
with the same layout (fixtures, exporting weights back).

PARITY UNPINNED: TensorFlow is not installable in the build environment, so this module is checked against the
published format (round trips, hand-assembled golden bytes, CRC-32C known answers), not against a file written
by TensorFlow itself.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57          # leveldb table footer magic, little-endian on disk
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5                     # 1 byte compression type + 4 bytes masked crc32c
RESTART_INTERVAL = 16

# tensorflow/core/framework/types.proto (the numeric types a Saver writes for this model)
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(ValueError):
    pass


# ---------------------------------------------------------------------------------------- CRC-32C
def _make_crc_table():
    poly = 0x82F63B78
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t.append(c)
    return t


_CRC_TABLE = _make_crc_table()


def _crc32c_scalar(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


# -- vectorised CRC for large tensors (the fc_1 matrix is hundreds of megabytes; a byte loop in Python is not an option)
_CRC_TABLE_NP = np.array(_CRC_TABLE, dtype=np.uint32)
_SEG = 1024          # bytes per lane


def _zero_operator(nbytes):
    """32x32 GF(2) matrix (32 uint32 columns) that advances a CRC register over `nbytes` zero bytes."""
    cols = [int(_CRC_TABLE[(1 << i) & 0xFF] ^ ((1 << i) >> 8)) for i in range(32)]      # one zero byte

    def mul(a, b):          # a o b: apply b, then a
        out = []
        for v in b:
            r, i = 0, 0
            while v:
                if v & 1:
                    r ^= a[i]
                v >>= 1
                i += 1
            out.append(r)
        return out
    result = [1 << i for i in range(32)]
    n = int(nbytes)
    while n:
        if n & 1:
            result = mul(cols, result)
        cols = mul(cols, cols)
        n >>= 1
    return np.array(result, dtype=np.uint32)


def _apply_operator(op, v):
    """op (32 columns) applied to an array of registers."""
    out = np.zeros_like(v)
    for i in range(32):
        out ^= np.where((v >> np.uint32(i)) & np.uint32(1), op[i], np.uint32(0)).astype(np.uint32)
    return out


def _crc32c_numpy(buf):
    """CRC-32C of a uint8 array: the buffer is cut into lanes of _SEG bytes whose CRCs advance together (one table
    lookup per byte position, vectorised over the lanes); lane CRCs are then folded pairwise with the zero-byte
    operator, crc(A || B) = Z_len(B)(crc(A)) ^ crc(B) (the identity behind zlib's crc32_combine)."""
    n = buf.size
    lanes = n // _SEG
    head = n - lanes * _SEG                       # a short leading piece keeps every lane the same length
    c0 = _crc32c_scalar(buf[:head].tobytes()) if head else None
    seg = buf[head:].reshape(lanes, _SEG)
    c = np.full(lanes, 0xFFFFFFFF, dtype=np.uint32)
    for j in range(_SEG):
        c = _CRC_TABLE_NP[(c ^ seg[:, j]) & np.uint32(0xFF)] ^ (c >> np.uint32(8))
    c ^= np.uint32(0xFFFFFFFF)
    # fold: after each level a lane covers twice the bytes; an odd lane out is carried in front (it is the leftmost)
    length = _SEG
    carry, carry_len = None, 0                    # CRC of a leading part not yet merged, and the bytes AFTER it so far
    while c.size > 1:
        if c.size & 1:
            first, c = int(c[0]), c[1:]
            if carry is None:
                carry, carry_len = first, 0
            else:                                  # carry || first
                carry = int(_apply_operator(_zero_operator(length), np.array([carry], np.uint32))[0]) ^ first
            # bytes that follow the carry are exactly the lanes left in c
        op = _zero_operator(length)
        c = _apply_operator(op, c[0::2]) ^ c[1::2]
        length *= 2
    total = int(c[0])
    if carry is not None:                          # carry || (everything folded so far)
        total = int(_apply_operator(_zero_operator(length), np.array([carry], np.uint32))[0]) ^ total
    if c0 is not None:
        total = int(_apply_operator(_zero_operator(n - head), np.array([c0], np.uint32))[0]) ^ total
    return total


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of leveldb tables and tensor bundles."""
    if crc == 0 and isinstance(data, np.ndarray) and data.dtype == np.uint8 and data.size >= (1 << 16):
        return _crc32c_numpy(np.ascontiguousarray(data))
    if crc == 0 and isinstance(data, (bytes, bytearray, memoryview)) and len(data) >= (1 << 16):
        return _crc32c_numpy(np.frombuffer(data, dtype=np.uint8))
    return _crc32c_scalar(data, crc)


def mask_crc(c):
    """leveldb / TF store crcs 'masked': rotate right by 15 and add a constant."""
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------- varints / protobuf
def _get_varint(buf, pos):
    r, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7
        if shift > 70:
            raise CheckpointError("varint too long")


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_message(buf):
    """Minimal protobuf wire decoder: [(field number, wire type, value)], value = int or bytes."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise CheckpointError("truncated length-delimited field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _field(tag, wt):
    return _put_varint((tag << 3) | wt)


def _parse_shape(buf):
    """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}; unknown_rank = 3."""
    dims = []
    for f, _, v in _parse_message(buf):
        if f == 2:
            size = 0
            for ff, _, vv in _parse_message(v):
                if ff == 1:
                    size = vv if vv < (1 << 63) else vv - (1 << 64)
            dims.append(size)
        elif f == 3 and v:
            raise CheckpointError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto)."""
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, wt, v in _parse_message(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _parse_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


def _parse_header(buf):
    """BundleHeaderProto: num_shards = 1, endianness = 2 (0 little), version = 3."""
    h = {"num_shards": 1, "endianness": 0}
    for f, _, v in _parse_message(buf):
        if f == 1:
            h["num_shards"] = v
        elif f == 2:
            h["endianness"] = v
    return h


# ---------------------------------------------------------------------------------------- snappy (raw format)
def _snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
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
            ln = ((tag >> 2) & 7) + 4
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
            raise CheckpointError("corrupt snappy block")
        for _ in range(ln):                             # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy length mismatch")
    return bytes(out)


# ---------------------------------------------------------------------------------------- table (SSTable) reader
def _read_block(data, offset, size, verify):
    end = offset + size
    if end + BLOCK_TRAILER_LEN > len(data):
        raise CheckpointError("table block handle out of range")
    body, ctype = data[offset:end], data[end]
    if verify:
        stored = struct.unpack_from("<I", data, end + 1)[0]
        if mask_crc(crc32c(data[offset:end + 1])) != stored:
            raise CheckpointError("table block checksum mismatch")
    if ctype == 0:
        return body
    if ctype == 1:
        return _snappy_uncompress(body)
    raise CheckpointError("unknown table block compression %d" % ctype)


def _block_entries(block):
    """Entries of one table block: prefix-compressed keys, restart array at the end."""
    if len(block) < 4:
        raise CheckpointError("table block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise CheckpointError("corrupt key prefix")
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < FOOTER_LEN:
        raise CheckpointError("%s: too short for a table footer" % path)
    footer = data[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, FOOTER_LEN - 8)[0] != TABLE_MAGIC:
        raise CheckpointError("%s: not a TensorFlow V2 checkpoint index (bad table magic)" % path)
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ---------------------------------------------------------------------------------------- public API
def list_variables(prefix):
    """[(name, shape, numpy dtype)] like tf.train.list_variables."""
    out = []
    for key, value in read_table(prefix + ".index"):
        if key == b"":
            continue
        e = _parse_entry(value)
        out.append((key.decode("utf-8"), e["shape"], _DTYPES.get(e["dtype"])))
    return out


def read_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} of a V2 checkpoint `prefix` (what tf.train.latest_checkpoint returns).
    `names`: optional subset.  `verify`: check table and tensor CRC-32C values."""
    entries = read_table(prefix + ".index", verify)
    if not entries or entries[0][0] != b"":
        raise CheckpointError("%s.index: missing bundle header entry" % prefix)
    header = _parse_header(entries[0][1])
    if header["endianness"] != 0:
        raise CheckpointError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    wanted = None if names is None else set(names)
    for key, value in entries[1:]:
        name = key.decode("utf-8")
        if wanted is not None and name not in wanted:
            continue
        e = _parse_entry(value)
        if e["slices"]:
            raise CheckpointError("%s: partitioned (sliced) variables are not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise CheckpointError("%s: unsupported dtype id %d" % (name, e["dtype"]))
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise CheckpointError("%s: size %d does not match shape %s of %s" % (name, e["size"], e["shape"], dt))
        sid = e["shard_id"]
        if sid not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"])
            if not os.path.exists(path):
                raise CheckpointError("missing data shard %s" % path)
            shards[sid] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise CheckpointError("%s: data shard is truncated" % name)
        buf = np.array(raw)          # the one copy out of the memory map (also the returned tensor's storage)
        if verify and e["crc32c"] is not None:
            if mask_crc(crc32c(buf)) != e["crc32c"]:
                raise CheckpointError("%s: tensor checksum mismatch" % name)
        out[name] = buf.view(dt).reshape(e["shape"])
    if wanted is not None and wanted - set(out):
        raise KeyError("not in checkpoint: %s" % sorted(wanted - set(out)))
    return out


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file of `directory`."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                p = line.split(":", 1)[1].strip().strip('"')
                return p if os.path.isabs(p) else os.path.join(directory, p)
    return None


# ---------------------------------------------------------------------------------------- writer
def _shape_proto(shape):
    out = b""
    for d in shape:
        dim = _field(1, 0) + _put_varint(int(d))
        out += _field(2, 2) + _put_varint(len(dim)) + dim
    return out


def _entry_proto(dtype_id, shape, offset, size, crc):
    sp = _shape_proto(shape)
    out = _field(1, 0) + _put_varint(dtype_id)
    out += _field(2, 2) + _put_varint(len(sp)) + sp
    if offset:
        out += _field(4, 0) + _put_varint(offset)
    out += _field(5, 0) + _put_varint(size)
    out += _field(6, 5) + struct.pack("<I", crc)
    return out


def _build_block(items):
    buf, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % RESTART_INTERVAL == 0:
            restarts.append(len(buf))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        buf += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def _emit_block(out, block):
    off = len(out)
    trailer = b"\x00"                                    # no compression
    out += block + trailer + struct.pack("<I", mask_crc(crc32c(block + trailer)))
    return off, len(block)


def write_checkpoint(prefix, variables, block_size=4096):
    """Write {name: ndarray} as a one-shard V2 checkpoint (`prefix.index`, `prefix.data-00000-of-00001`)
    plus the `checkpoint` state file next to it."""
    names = sorted(variables, key=lambda s: s.encode("utf-8"))
    data = bytearray()
    header = _field(1, 0) + _put_varint(1) + _field(3, 2) + _put_varint(2) + _field(1, 0) + _put_varint(1)
    items = [(b"", header)]
    for n in names:
        a = np.ascontiguousarray(variables[n])
        if a.dtype not in _DTYPE_IDS:
            raise CheckpointError("%s: dtype %s cannot be written" % (n, a.dtype))
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        items.append((n.encode("utf-8"), _entry_proto(_DTYPE_IDS[a.dtype], a.shape, len(data), len(raw),
                                                      mask_crc(crc32c(raw)))))
        data += raw
    out = bytearray()
    index_items, cur, cur_bytes = [], [], 0
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 6
        if cur_bytes >= block_size:
            off, size = _emit_block(out, _build_block(cur))
            index_items.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
            cur, cur_bytes = [], 0
    if cur:
        off, size = _emit_block(out, _build_block(cur))
        index_items.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
    meta_off, meta_size = _emit_block(out, _build_block([]))
    idx_off, idx_size = _emit_block(out, _build_block(index_items))
    footer = _put_varint(meta_off) + _put_varint(meta_size) + _put_varint(idx_off) + _put_varint(idx_size)
    footer += b"\x00" * (FOOTER_LEN - 8 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(out)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
