"""Test helper: writes a TensorFlow TensorBundle (`<prefix>.index`, `<prefix>.data-00000-of-00001`) from the published
layout -- LevelDB-style table with prefix-compressed keys, restart points, several data blocks, block trailers with masked
CRC-32C, BundleHeaderProto / BundleEntryProto values.  Written independently of the reader in the package (it shares
only the CRC helper), so that tests exercise the reader on multi-block tables; it is NOT TensorFlow's own writer."""
import struct

import numpy as np

from neural_network_image_compression_b200.tfbundle import TABLE_MAGIC, crc32c, mask_crc

DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def field(num, wt, payload):
    return varint((num << 3) | wt) + payload


def snappy_compress(data: bytes) -> bytes:
    """A small greedy Snappy compressor (4-byte hash matches, every element kind): enough to write blocks a decoder has to
    take apart properly -- literals of all length encodings, copies with 1- and 2-byte offsets, overlapping copies."""
    out = bytearray(varint(len(data)))
    table, pos, lit = {}, 0, 0

    def flush_literal(end):
        nonlocal lit
        while lit < end:
            n = min(end - lit, 70000)
            if n <= 60:
                out.append((n - 1) << 2)
            elif n <= 256:
                out.extend(bytes([60 << 2, n - 1]))
            else:
                nb = 2 if n <= 65536 else 3
                out.append((59 + nb) << 2)
                out.extend((n - 1).to_bytes(nb, "little"))
            out.extend(data[lit:lit + n])
            lit += n
    while pos + 4 <= len(data):
        key = data[pos:pos + 4]
        cand = table.get(key)
        table[key] = pos
        if cand is not None and pos - cand <= 65535:
            ln = 4
            while pos + ln < len(data) and ln < 64 and data[cand + ln] == data[pos + ln]:
                ln += 1
            flush_literal(pos)
            off = pos - cand
            if 4 <= ln <= 11 and off < 2048:
                out.extend(bytes([1 | ((ln - 4) << 2) | ((off >> 8) << 5), off & 0xFF]))
            else:
                out.extend(bytes([2 | ((ln - 1) << 2), off & 0xFF, off >> 8]))
            pos += ln
            lit = pos
        else:
            pos += 1
    flush_literal(len(data))
    return bytes(out)


def entry_proto(arr, offset, shard=0):
    shape = b"".join(field(2, 2, varint(len(d)) + d) for d in (field(1, 0, varint(s)) for s in arr.shape))
    raw = arr.tobytes()
    return (field(1, 0, varint(DT[arr.dtype])) + field(2, 2, varint(len(shape)) + shape)
            + (field(3, 0, varint(shard)) if shard else b"") + field(4, 0, varint(offset))
            + field(5, 0, varint(len(raw))) + field(6, 5, struct.pack("<I", mask_crc(crc32c(raw)))))


class BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_bundle(prefix, tensors, block_bytes=300, string_entries=("_CHECKPOINTABLE_OBJECT_GRAPH",), num_shards=1, snappy=False):
    """tensors: {key: ndarray}.  Keys are written in sorted order; `block_bytes` small => many data blocks;
    `num_shards` > 1 spreads the tensors round-robin over `.data-NNNNN-of-MMMMM` files; `snappy` compresses every table block."""
    datas, entries = [bytearray() for _ in range(num_shards)], {}
    header = field(1, 0, varint(num_shards)) + field(2, 0, varint(0)) + field(3, 2, varint(2) + field(1, 0, varint(1)))
    entries[b""] = header
    for i, key in enumerate(sorted(tensors)):
        arr = np.ascontiguousarray(tensors[key])
        shard = i % num_shards
        entries[key.encode()] = entry_proto(arr, len(datas[shard]), shard)
        datas[shard] += arr.tobytes()
    for key in string_entries:                       # a DT_STRING entry, as Keras writes for the object graph
        blob = b"\x05graph"
        entries[key.encode()] = (field(1, 0, varint(7)) + field(2, 2, varint(0)) + field(4, 0, varint(len(datas[0])))
                                 + field(5, 0, varint(len(blob))))
        datas[0] += blob
    for shard, data in enumerate(datas):
        with open(f"{prefix}.data-{shard:05d}-of-{num_shards:05d}", "wb") as f:
            f.write(data)

    out = bytearray()

    def emit(block):
        off = len(out)
        kind = b"\x01" if snappy else b"\x00"
        stored = snappy_compress(block) if snappy else block
        out.extend(stored + kind + struct.pack("<I", mask_crc(crc32c(stored + kind))))
        return varint(off) + varint(len(stored))

    index = BlockBuilder(restart_interval=1)
    cur = BlockBuilder()
    for key in sorted(entries):
        cur.add(key, entries[key])
        if len(cur.buf) >= block_bytes:
            index.add(cur.last, emit(cur.finish()))
            cur = BlockBuilder()
    if cur.count:
        index.add(cur.last, emit(cur.finish()))
    meta_handle = emit(BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(f"{prefix}.index", "wb") as f:
        f.write(out)
