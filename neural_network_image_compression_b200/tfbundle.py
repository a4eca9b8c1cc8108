"""Reader for TensorFlow "TensorBundle" checkpoints (`<prefix>.index` + `<prefix>.data-00000-of-00001`), the format
`tf.keras.Model.save_weights(prefix)` writes and `load_weights(prefix)` reads in the reference
(/root/reference/tf2_0/src/utils.py:26-28, training.py:167-170).

No TensorFlow is involved: the index file is a LevelDB-style sorted string table, its values are `BundleEntryProto`
protobufs, the data file holds the raw little-endian tensors.  What is implemented here is the published layout:

  table file   = data blocks, meta-index block, index block, 48-byte footer
  footer       = BlockHandle(meta-index) BlockHandle(index), zero padding to 40 bytes, magic 0xdb4775248b80fb57 (LE)
  BlockHandle  = varint64 offset, varint64 size           (size excludes the 5-byte block trailer: type + crc32c)
  block        = entries, uint32 restart offsets, uint32 number of restarts
  entry        = varint32 shared key bytes, varint32 unshared key bytes, varint32 value bytes, key suffix, value
  key ""       -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
  other keys   -> BundleEntryProto {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset, 5: size,
                                    6: crc32c (masked, fixed32)}

STATUS: this container has neither TensorFlow nor a checkpoint of the reference, so the reader is exercised against
files produced by an independently written writer of the same layout (tests/tf_bundle_writer.py), not against
TensorFlow's own output.  Snappy-compressed table blocks (LevelDB block type 1; TensorFlow's own writer leaves them
uncompressed) and checkpoints sharded over several `.data-NNNNN-of-MMMMM` files are read; sliced tensors and dtypes other
than the ones listed in _DTYPES are rejected with an error instead of being guessed at.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16}   # tensorflow DataType enum
VARIABLE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


# ---- CRC-32C (Castagnoli), with TensorFlow's masking --------------------------------------------------------------
def _crc_table():
    tab = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab[i] = c
    return [int(v) for v in tab]


_CRC_TABLE = _crc_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints and the few protobuf fields needed ------------------------------------------------------------------------
def _varint(buf: bytes, pos: int):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _proto_fields(buf: bytes):
    """Yield (field number, wire type, value) of one protobuf message; length-delimited values come as bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val = bytes(buf[pos:pos + n]); pos += n
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, val


def _parse_entry(value: bytes):
    dtype = shard = offset = size = 0
    crc = None
    shape = []
    for field, _wt, val in _proto_fields(value):
        if field == 1:
            dtype = val
        elif field == 2:
            for f2, _w2, v2 in _proto_fields(val):
                if f2 == 2:                               # Dim
                    dim = 0
                    for f3, _w3, v3 in _proto_fields(v2):
                        if f3 == 1:
                            dim = v3
                    shape.append(dim)
                elif f2 == 3 and v2:
                    raise ValueError("tensor of unknown rank")
        elif field == 3:
            shard = val
        elif field == 4:
            offset = val
        elif field == 5:
            size = val
        elif field == 6:
            crc = val
        elif field == 7:
            raise ValueError("sliced tensors are not supported")
    return dtype, tuple(shape), shard, offset, size, crc


# ---- the sorted string table -------------------------------------------------------------------------------------
def _block_handle(buf: bytes, pos: int):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def snappy_uncompress(buf: bytes) -> bytes:
    """Raw Snappy block format (the compression LevelDB tables use): varint uncompressed length, then literal and copy
    elements.  tag & 3: 0 literal (length - 1 in the upper six bits, 60..63 = that many - 59 extra length bytes),
    1 copy with an 11-bit offset (length 4..11), 2 copy with a 16-bit offset, 3 copy with a 32-bit offset (length 1..64)."""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                extra = ln - 59
                ln = int.from_bytes(buf[pos:pos + extra], "little"); pos += extra
            ln += 1
            if pos + ln > len(buf):
                raise ValueError("snappy: literal runs past the end of the block")
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            offset = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            offset = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            offset = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        if offset == 0 or offset > len(out):
            raise ValueError("snappy: copy offset outside the data written so far")
        for _ in range(ln):                      # byte by byte: copies may overlap their own output (run-length encoding)
            out.append(out[-offset])
    if len(out) != n:
        raise ValueError(f"snappy: block expands to {len(out)} bytes, header says {n}")
    return bytes(out)


def _read_block(data: bytes, off: int, size: int, verify: bool) -> bytes:
    body, trailer = data[off:off + size], data[off + size:off + size + 5]
    if len(body) != size or len(trailer) != 5:
        raise ValueError("block handle points outside the index file")
    if trailer[0] not in (0, 1):
        raise ValueError("unknown table block type %d (0 = raw, 1 = snappy)" % trailer[0])
    if verify:                                   # the checksum covers the stored (possibly compressed) bytes + type
        want = struct.unpack("<I", trailer[1:])[0]
        if mask_crc(crc32c(body + trailer[:1])) != want:
            raise ValueError("index block checksum mismatch")
    return snappy_uncompress(body) if trailer[0] == 1 else body


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 * (n_restarts + 1)
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str, verify: bool = True) -> dict:
    """`<prefix>.index` -> {key (str): raw value bytes}, keys in file order (sorted)."""
    with open(index_path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError(f"{index_path} is not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _mo, _ms, pos = _block_handle(footer, 0)
    io, isz, _ = _block_handle(footer, pos)
    out = {}
    for _last_key, handle in _block_entries(_read_block(data, io, isz, verify)):
        off, size, _ = _block_handle(handle, 0)
        for key, value in _block_entries(_read_block(data, off, size, verify)):
            out[key.decode("utf-8")] = bytes(value)
    return out


def read_bundle(prefix: str, verify: bool = True) -> dict:
    """All numeric tensors of the checkpoint `prefix` as {key: ndarray}.  String tensors (the object graph) are skipped."""
    index = read_index(prefix + ".index", verify)
    num_shards = 1
    for field, _wt, val in _proto_fields(index.get("", b"")):
        if field == 1:
            num_shards = val
        elif field == 2 and val != 0:
            raise ValueError("big-endian checkpoints are not supported")
    if num_shards < 1:
        raise ValueError(f"checkpoint header names {num_shards} shards")
    files = {}
    out = {}
    try:
        for key, value in index.items():
            if key == "":
                continue
            dtype, shape, shard, offset, size, crc = _parse_entry(value)
            if dtype == 7:                                   # DT_STRING: _CHECKPOINTABLE_OBJECT_GRAPH
                continue
            if dtype not in _DTYPES:
                raise ValueError(f"{key}: unsupported dtype {dtype}")
            if not 0 <= shard < num_shards:
                raise ValueError(f"{key}: shard {shard} of a checkpoint with {num_shards} shards")
            if shard not in files:
                files[shard] = open(f"{prefix}.data-{shard:05d}-of-{num_shards:05d}", "rb")
            f = files[shard]
            f.seek(offset)
            raw = f.read(size)
            n = int(np.prod(shape, dtype=np.int64)) if shape else 1
            if len(raw) != size or size != n * np.dtype(_DTYPES[dtype]).itemsize:
                raise ValueError(f"{key}: data file too short or size does not match shape {shape}")
            if verify and crc is not None and mask_crc(crc32c(raw)) != crc:
                raise ValueError(f"{key}: tensor checksum mismatch")
            out[key] = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dtype]).newbyteorder("<")).reshape(shape).copy()
    finally:
        for f in files.values():
            f.close()
    return out


def keras_weights(prefix: str, layer_names, verify: bool = True) -> dict:
    """Weights of one network saved by `model.save_weights(prefix)` as {'<layer>/kernel', '<layer>/bias'}.

    Keras' object-based checkpoints name a variable by its path from the model object:
    `<attribute>/kernel/.ATTRIBUTES/VARIABLE_VALUE` (the reference's models keep their layers in attributes conv1 ...
    dconv8, encoder.py:10-17) or, for layers tracked by position, `layer_with_weights-<i>/kernel/...`; both are accepted."""
    tensors = read_bundle(prefix, verify)
    out = {}
    for i, name in enumerate(layer_names):
        for var in ("kernel", "bias"):
            for key in (f"{name}/{var}{VARIABLE_SUFFIX}", f"layer_with_weights-{i}/{var}{VARIABLE_SUFFIX}"):
                if key in tensors:
                    out[f"{name}/{var}"] = np.asarray(tensors[key], np.float32)
                    break
            else:
                raise KeyError(f"{prefix}: no variable for {name}/{var} (keys: {sorted(tensors)[:6]} ...)")
    return out


def is_bundle(prefix: str) -> bool:
    return os.path.exists(prefix + ".index")
