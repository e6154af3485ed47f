"""Standalone reader (and writer) for TF2 object-based checkpoints -- SURVEY 8(f) row 1.

The reference restores its three models with `tf.train.Checkpoint(unet=unet).restore(path)`
(run_ldm_sampler.py:70-75) from the files `convert_ckpt_pytorch_to_tf2.py:426-431` wrote:
`<prefix>.index` + `<prefix>.data-00000-of-00001`.  TensorFlow is not installable here, so this
module restates the published TensorBundle format (tensorflow/core/util/tensor_bundle, TF 2.13,
not vendored by the reference) in plain Python:

  * `.index` is a LevelDB-style sorted string table (tensorflow/core/lib/io/table): data blocks of
    prefix-compressed entries `[shared varint32][non_shared varint32][value_len varint32][key
    suffix][value]` followed by a restart array (`uint32` offsets, `uint32` count); every block is
    followed by a 5-byte trailer (compression type, masked CRC-32C); the last 48 bytes are the
    footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).  The bundle writer
    uses no compression.
  * key "" holds a `BundleHeaderProto` (num_shards, endianness, version); every other key holds a
    `BundleEntryProto` (dtype, shape, shard_id, offset, size, masked crc32c of the payload).
  * `.data-NNNNN-of-MMMMM` holds the raw little-endian tensor bytes at (offset, size).
  * variable keys are `<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`; the library
    reports exactly these attribute paths as weight names (`ldm_weight_info`), in flat Keras order,
    which tests/golden/ckpt_keys_small.json pins against the reference's own objects.

Parity status: no real TensorFlow checkpoint exists in this environment, so the reader is verified
against this module's writer (both follow the format description above), the CRC-32C known answer,
and hand-built tables with prefix compression; reading a file written by TensorFlow itself is
UNPINNED.  The writer emits name-based bundles (`tf.train.load_checkpoint` / this reader); it does not
write the `_CHECKPOINTABLE_OBJECT_GRAPH` proto `Checkpoint.restore` needs.
"""
import ctypes as C
import os
import struct

import numpy as np

SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_STRING, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 1, 2, 3, 4, 7, 9, 10, 14, 19
_NP_OF_DT = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_UINT8: np.uint8, DT_INT64: np.int64,
             DT_BOOL: np.bool_, DT_HALF: np.float16}
_DT_OF_NP = {np.dtype(v): k for k, v in _NP_OF_DT.items()}


class CheckpointError(RuntimeError):
    pass


# ----------------------------------------------------------------------------- CRC-32C
_TABLE = None


def _crc32c_py(data: bytes, crc: int = 0) -> int:
    global _TABLE
    if _TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _TABLE = t
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C (Castagnoli).  Large buffers go through libldm_b200's SSE4.2 routine when the library
    is built; small ones (index blocks) through the table above."""
    mv = memoryview(data).cast("B")
    if len(mv) >= 4096:
        try:
            from . import lib as _lib
            L = _lib.load()
            out = C.c_uint()
            arr = np.frombuffer(mv, dtype=np.uint8)
            if L.ldm_crc32c(arr.ctypes.data_as(C.c_void_p), arr.size, crc, C.byref(out)) == 0:
                return out.value
        except Exception:
            pass
    return _crc32c_py(bytes(mv), crc)


def mask_crc(c: int) -> int:
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varints / protobuf
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


def _pb_fields(buf):
    """Yields (field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, w = tag >> 3, tag & 7
        if w == 0:
            v, pos = _get_varint(buf, pos)
        elif w == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif w == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise CheckpointError("truncated protobuf field")
            pos += n
        elif w == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError(f"unsupported protobuf wire type {w}")
        yield f, w, v


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard, offset, size, crc, sliced)."""
    e = dict(dtype=0, shape=(), shard=0, offset=0, size=0, crc=0, sliced=False)
    for f, _, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:  # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1; } bool unknown_rank = 3 }
            dims = []
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif f == 3:
            e["shard"] = v
        elif f == 4:
            e["offset"] = _signed64(v)
        elif f == 5:
            e["size"] = _signed64(v)
        elif f == 6:
            e["crc"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def _build_entry(dtype, shape, offset, size, crc):
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ----------------------------------------------------------------------------- table (.index)
def _read_block(buf, offset, size, verify):
    if offset + size + 5 > len(buf):
        raise CheckpointError("block handle outside the index file")
    body = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and mask_crc(crc32c(buf[offset:offset + size + 1])) != stored:
        raise CheckpointError("index block checksum mismatch")
    if ctype != 0:
        raise CheckpointError("compressed index blocks are not supported (the bundle writer never compresses)")
    return body


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("index block too short")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    if end < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise CheckpointError("corrupt index entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(prefix, verify=True):
    """`<prefix>.index` -> (header dict, {key: entry dict}) in key order."""
    path = prefix + ".index"
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48:
        raise CheckpointError(f"{path}: too short for a table footer")
    footer = buf[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != MAGIC:
        raise CheckpointError(f"{path}: bad table magic (not a TensorBundle index)")
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        for key, val in _block_entries(_read_block(buf, boff, bsize, verify)):
            if key == b"":
                header = dict(num_shards=1, endianness=0)
                for f, _, v in _pb_fields(val):
                    if f == 1:
                        header["num_shards"] = v
                    elif f == 2:
                        header["endianness"] = v
            else:
                entries[key.decode()] = _parse_entry(val)
    if header is None:
        raise CheckpointError(f"{path}: no bundle header entry")
    if header["endianness"] != 0:
        raise CheckpointError("big-endian bundles are not supported")
    return header, entries


def load_checkpoint(prefix, keys=None, verify=True):
    """Reads tensors of a TF2 checkpoint into NumPy arrays: {key: array}.  `keys` restricts the read
    (missing ones raise); string-typed entries (the object graph) are skipped."""
    header, entries = read_index(prefix, verify)
    want = list(entries) if keys is None else list(keys)
    files, out = {}, {}
    try:
        for k in want:
            if k not in entries:
                raise CheckpointError(f"{prefix}: no tensor named {k!r}")
            e = entries[k]
            if e["dtype"] == DT_STRING or e["sliced"]:
                if keys is not None:
                    raise CheckpointError(f"{k}: string / sliced tensors are not supported")
                continue
            if e["dtype"] not in _NP_OF_DT:
                raise CheckpointError(f"{k}: unsupported dtype enum {e['dtype']}")
            dt = np.dtype(_NP_OF_DT[e["dtype"]])
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise CheckpointError(f"{k}: shape {e['shape']} does not match {e['size']} bytes")
            if e["shard"] not in files:
                name = f"{prefix}.data-{e['shard']:05d}-of-{header['num_shards']:05d}"
                files[e["shard"]] = np.memmap(name, dtype=np.uint8, mode="r") if os.path.getsize(name) else np.zeros(0, np.uint8)
            raw = files[e["shard"]][e["offset"]:e["offset"] + e["size"]]
            if raw.size != e["size"]:
                raise CheckpointError(f"{k}: data file too short")
            if verify and mask_crc(crc32c(raw)) != e["crc"]:
                raise CheckpointError(f"{k}: payload checksum mismatch")
            out[k] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e["shape"])
    finally:
        files.clear()
    return out


def _table_block(items, restart_interval):
    body, restarts, prev, n = bytearray(), [], b"", 0
    for key, val in items:
        if n % restart_interval == 0:
            restarts.append(len(body))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        prev = key
        n += 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_checkpoint(prefix, tensors, block_size=4096, restart_interval=16):
    """Writes {key: array} as a single-shard TensorBundle (`<prefix>.index`, `.data-00000-of-00001`)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", b"\x08\x01\x1a\x02\x08\x01")]   # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for key in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[key])
            if a.dtype not in _DT_OF_NP:
                raise CheckpointError(f"{key}: dtype {a.dtype} not supported")
            raw = a.tobytes()   # C order, little endian
            f.write(raw)
            items.append((key.encode(), _build_entry(_DT_OF_NP[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    out = bytearray()

    def emit(block):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block)
        out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return handle

    index_items, cur, cur_bytes = [], [], 0
    for kv in items:
        cur.append(kv)
        cur_bytes += len(kv[0]) + len(kv[1]) + 3
        if cur_bytes >= block_size:
            index_items.append((cur[-1][0], emit(_table_block(cur, restart_interval))))
            cur, cur_bytes = [], 0
    if cur:
        index_items.append((cur[-1][0], emit(_table_block(cur, restart_interval))))
    meta = emit(_table_block([], 1))
    index = emit(_table_block(index_items, 1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


# ----------------------------------------------------------------------------- model glue
def variable_keys(handle, model):
    """Checkpoint keys of one model's variables in flat Keras order (works on a describe-only handle)."""
    return [handle.weight_info(model, i)[0] + SUFFIX for i in range(handle.num_weights(model))]


def restore(handle, model, prefix, verify=True):
    """tf.train.Checkpoint(<root>=layer).restore(prefix) for one of the three models
    (run_ldm_sampler.py:70-75): reads every variable by its attribute-path key, validates shapes and
    hands the flat list to the library."""
    keys = variable_keys(handle, model)
    tensors = load_checkpoint(prefix, keys, verify)
    weights = []
    for i, k in enumerate(keys):
        _, shape = handle.weight_info(model, i)
        a = tensors[k]
        if a.dtype != np.float32 or tuple(a.shape) != tuple(shape):
            raise CheckpointError(f"{k}: checkpoint has {a.dtype}{tuple(a.shape)}, model expects float32{tuple(shape)}")
        weights.append(a)
    handle.set_weights(model, weights)
    return len(weights)


def save(handle_or_names, weights, prefix, model=None):
    """Writes a flat Keras weight list under the reference's checkpoint keys (name-based bundle)."""
    keys = variable_keys(handle_or_names, model) if model is not None else list(handle_or_names)
    if len(keys) != len(weights):
        raise CheckpointError(f"{len(keys)} keys for {len(weights)} tensors")
    write_checkpoint(prefix, {k: np.asarray(w, np.float32) for k, w in zip(keys, weights)})
