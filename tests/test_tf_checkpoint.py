"""TensorFlow-free V2 checkpoint reader (monkey_pose_b200.tf_checkpoint): format conformance on CPU.

TensorFlow cannot be installed here, so no file written by TF itself is available ("parity unpinned" for this
module).  What pins it instead: CRC-32C known-answer vectors (RFC 3720), a bundle whose bytes are assembled in
this file by an independent, deliberately naive encoder that follows the published formats (leveldb
table_format.md, tensor_bundle.proto), a snappy-compressed block, corruption detection and round trips."""
import os
import struct

import numpy as np
import pytest

import monkey_pose_b200  # noqa: F401  (path alias)
from monkey_pose_b200 import tf_checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 B.4 / the leveldb crc32c_test vectors
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert ck.crc32c(b"123456789") == 0xE3069283
    # masking is a bijection that changes the value (leveldb crc32c.h)
    c = ck.crc32c(b"foo")
    assert ck.mask_crc(c) != c and ck.mask_crc(c) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- an independent, naive encoder of the two formats (no code shared with the module) ----------------------
def _vi(v):
    b = b""
    while v >= 128:
        b += bytes([v & 127 | 128])
        v >>= 7
    return b + bytes([v])


def _naive_entry(dtype_id, shape, offset, size, crc):
    shape_pb = b"".join(b"\x12" + _vi(len(b"\x08" + _vi(d))) + b"\x08" + _vi(d) for d in shape)
    pb = b"\x08" + _vi(dtype_id) + b"\x12" + _vi(len(shape_pb)) + shape_pb
    pb += b"\x20" + _vi(offset) + b"\x28" + _vi(size) + b"\x35" + struct.pack("<I", crc)
    return pb


def _naive_block(items, compress=None):
    """No key prefix sharing at all (every entry: shared = 0), a single restart point."""
    body = b"".join(_vi(0) + _vi(len(k)) + _vi(len(v)) + k + v for k, v in items)
    body += struct.pack("<I", 0) + struct.pack("<I", 1)
    ctype = b"\x00"
    if compress:
        body, ctype = compress(body), b"\x01"
    crc = ck.mask_crc(ck.crc32c(body + ctype))
    return body + ctype + struct.pack("<I", crc), len(body)


def _naive_bundle(tmp_path, tensors, compress=None, name="golden.ckpt"):
    data, items = b"", [(b"", b"\x08\x01\x1a\x02\x08\x01")]      # header: num_shards 1, version{producer 1}
    for n in sorted(tensors):
        a = tensors[n]
        raw = a.tobytes()
        dt = {np.dtype(np.float32): 1, np.dtype(np.int64): 9, np.dtype(np.int32): 3}[a.dtype]
        items.append((n.encode(), _naive_entry(dt, a.shape, len(data), len(raw), ck.mask_crc(ck.crc32c(raw)))))
        data += raw
    out = b""
    blk, blen = _naive_block(items, compress)
    data_handle = _vi(0) + _vi(blen)
    out += blk
    meta_off = len(out)
    blk, mlen = _naive_block([])
    out += blk
    idx_off = len(out)
    blk, ilen = _naive_block([(items[-1][0] + b"\xff", data_handle)])
    out += blk
    footer = _vi(meta_off) + _vi(mlen) + _vi(idx_off) + _vi(ilen)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    prefix = str(tmp_path / name)
    with open(prefix + ".index", "wb") as f:
        f.write(out + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)
    return prefix


def _tensors():
    rng = np.random.default_rng(0)
    return {"cnn/contextual_circuit/p_r": rng.standard_normal((3, 3, 2, 2)).astype(np.float32),
            "cnn/conv_1/conv_1_biases": rng.standard_normal((5,)).astype(np.float32),
            "cnn/conv_1/conv_1_biases/Adam": np.zeros((5,), np.float32),
            "global_step": np.array(1234, np.int64)}


def test_reads_hand_assembled_bundle(tmp_path):
    t = _tensors()
    prefix = _naive_bundle(tmp_path, t)
    got = ck.read_checkpoint(prefix)
    assert sorted(got) == sorted(t)
    for n in t:
        assert got[n].dtype == t[n].dtype and got[n].shape == t[n].shape
        np.testing.assert_array_equal(got[n], t[n])
    listed = dict((n, (s, d)) for n, s, d in ck.list_variables(prefix))
    assert listed["cnn/contextual_circuit/p_r"] == ((3, 3, 2, 2), np.float32)
    assert listed["global_step"] == ((), np.int64)
    sub = ck.read_checkpoint(prefix, names=["global_step"])
    assert list(sub) == ["global_step"] and int(sub["global_step"]) == 1234
    with pytest.raises(KeyError):
        ck.read_checkpoint(prefix, names=["nope"])


def _snappy_literal_and_copies(raw):
    """A valid raw-snappy stream: literals in <= 60-byte runs, plus one back-reference (copy, 2-byte offset)
    wherever the next 8 bytes repeat the 8 bytes before them -- exercises both element kinds."""
    out, pos = _vi(len(raw)), 0
    while pos < len(raw):
        if pos >= 8 and raw[pos:pos + 8] == raw[pos - 8:pos] and len(raw) - pos >= 8:
            out += bytes([((8 - 1) << 2) | 2]) + struct.pack("<H", 8)
            pos += 8
            continue
        n = min(60, len(raw) - pos)
        out += bytes([(n - 1) << 2]) + raw[pos:pos + n]
        pos += n
    return out


def test_reads_snappy_compressed_table_blocks(tmp_path):
    t = {"a/zeros": np.zeros((64,), np.float32), "b/ramp": np.arange(7, dtype=np.int32)}
    prefix = _naive_bundle(tmp_path, t, compress=_snappy_literal_and_copies, name="snappy.ckpt")
    got = ck.read_checkpoint(prefix)
    for n in t:
        np.testing.assert_array_equal(got[n], t[n])
    assert ck._snappy_uncompress(_snappy_literal_and_copies(b"abcdefgh" * 5 + b"xyz")) == b"abcdefgh" * 5 + b"xyz"


def test_detects_corruption(tmp_path):
    prefix = _naive_bundle(tmp_path, _tensors())
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[3] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(ck.CheckpointError, match="checksum"):
        ck.read_checkpoint(prefix)
    assert ck.read_checkpoint(prefix, verify=False)       # readable when verification is off
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ck.CheckpointError, match="checksum"):
        ck.read_table(prefix + ".index")
    open(prefix + ".index", "wb").write(b"not a table" * 10)
    with pytest.raises(ck.CheckpointError, match="magic"):
        ck.read_table(prefix + ".index")


def test_write_read_round_trip_many_blocks_and_latest_checkpoint(tmp_path):
    rng = np.random.default_rng(1)
    # > 16 keys with long shared prefixes: prefix compression, restart points and several data blocks
    t = {"cnn/batch_normalization_%d/moving_variance" % i: rng.random((4 + i,)).astype(np.float32)
         for i in range(40)}
    t["cnn/fc_1/fc_1_weights"] = rng.standard_normal((64, 8)).astype(np.float32)
    t["step"] = np.array(7, np.int64)
    prefix = str(tmp_path / "model_99.ckpt-99")
    ck.write_checkpoint(prefix, t, block_size=256)
    assert ck.latest_checkpoint(str(tmp_path)) == prefix
    got = ck.read_checkpoint(prefix)
    assert sorted(got) == sorted(t)
    for n in t:
        np.testing.assert_array_equal(got[n], t[n])
    keys = [k for k, _ in ck.read_table(prefix + ".index")]
    assert keys == sorted(keys) and keys[0] == b""


def test_model_load_checkpoint_maps_reference_names(tmp_path):
    """model.load_checkpoint: 'cnn/' scope stripped, Adam slots and global_step ignored, data_dict filled with the
    reference's variable names (hgru_pose.py:196-216)."""
    import monkey_pose_b200 as mp
    from monkey_pose_b200 import initialization as init
    P = init.pose_params(channels=3, S=5, T=2, hw=8, fc_hidden=16, out=6, seed=4)
    flat = {"cnn/" + k: np.asarray(v) for k, v in P.items()}
    flat["cnn/conv_1/conv_1_filters/Adam"] = np.zeros_like(flat["cnn/conv_1/conv_1_filters"])
    flat["cnn/conv_1/conv_1_filters/Adam_1"] = np.zeros_like(flat["cnn/conv_1/conv_1_filters"])
    flat["beta1_power"] = np.array(0.9, np.float32)
    flat["global_step"] = np.array(3, np.int64)
    flat["attn/aconv_1/aconv_1_filters"] = np.zeros((3, 3, 1, 4), np.float32)     # another network's scope
    prefix = str(tmp_path / "best.ckpt-3")
    ck.write_checkpoint(prefix, flat)
    m = mp.model()
    loaded = m.load_checkpoint(str(tmp_path))              # directory -> latest_checkpoint
    assert "conv_1/conv_1_filters" in loaded and not any("Adam" in n or n.startswith("attn") for n in loaded)
    np.testing.assert_array_equal(m.data_dict["conv_1"][0], P["conv_1/conv_1_filters"])
    np.testing.assert_array_equal(m.data_dict["contextual_circuit"]["p_r"], P["contextual_circuit/p_r"])
    np.testing.assert_array_equal(m.data_dict["fc_out"][1], P["fc_out/fc_out_biases"])
    os.remove(prefix + ".data-00000-of-00001")
    with pytest.raises(ck.CheckpointError, match="missing data shard"):
        mp.model().load_checkpoint(prefix)


def test_model_load_checkpoint_from_the_combined_graph(tmp_path):
    """Checkpoints of the reference's training graph hold BOTH networks under 'cnn/' (train_cnn_networks_hgru.py:
    96-142): the attention CNN's batch-norm layers are batch_normalization .. _5, the pose model's _6 .. _10."""
    import monkey_pose_b200 as mp
    from monkey_pose_b200 import initialization as init
    P = init.pose_params(channels=3, S=5, T=2, hw=8, fc_hidden=16, out=6, seed=4, random_bn=True)
    A = init.attn_params(widths=(8, 8, 8, 8, 8), fc_hidden=16, out=3, seed=2, random_bn=True)
    flat = {"cnn/" + k: np.asarray(v) for k, v in A.items()}
    for k, v in P.items():
        if k.startswith("batch_normalization"):
            scope, field = k.split("/")
            i = 0 if scope == "batch_normalization" else int(scope.rsplit("_", 1)[1])
            k = "batch_normalization_%d/%s" % (i + 6, field)
        flat["cnn/" + k] = np.asarray(v)
    prefix = str(tmp_path / "both.ckpt-1")
    ck.write_checkpoint(prefix, flat)
    m = mp.model()
    m.load_checkpoint(prefix)
    np.testing.assert_array_equal(m.data_dict["batch_normalization_3"][2], P["batch_normalization_3/moving_mean"])
    np.testing.assert_array_equal(m.data_dict["batch_normalization"][0], P["batch_normalization/gamma"])
    am = mp.attn_model_struct()
    am.load_checkpoint(prefix)
    np.testing.assert_array_equal(am.data_dict["batch_normalization_5"][3], A["batch_normalization_5/moving_variance"])
    np.testing.assert_array_equal(am.data_dict["aconv_5"][0], A["aconv_5/aconv_5_filters"])
