"""TF2 checkpoint (TensorBundle) reader / writer -- SURVEY 8(f) row 1.  CPU-only tests.

No TensorFlow-written checkpoint exists in this environment (parity against real TF is unpinned);
what is pinned: the CRC-32C known answers, the table format against a hand-assembled index with
prefix-compressed keys and several data blocks, corruption detection, the variable keys against
the reference's own objects (tests/golden/ckpt_keys_small.json, made by make_ckpt_keys.py), and -- against
TensorFlow's own code as shipped in tensorboard.compat -- the masked CRC-32C routine, the dtype enum, the
TensorShapeProto encoding and the TrackableObjectGraph parser / builder (byte-identical to TF's serializer)."""
import json
import os
import struct

import numpy as np
import pytest

from ldm_tf2_b200 import lib, tf_checkpoint as T

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = {
    "cond_stage_model": dict(vocab_size=30522, encoder_stack_size=2, hidden_size=1280, num_heads=8, size_per_head=64,
                             max_seq_len=77, filter_size=512),
    "unet": dict(model_channels=160, out_channels=4, num_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=4,
                 head_base=40, context_dim=1280),
    "autoencoder_kl": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[],
                           multipliers=[1, 2, 4, 4]),
    "autoencoder_vq": dict(latent_channels=4, channels=32, num_blocks=2, attention_resolutions=[8],
                           multipliers=[1, 2, 2, 4], vocab_size=512),
}


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert T._crc32c_py(b"123456789") == 0xE3069283
    assert T._crc32c_py(bytes(32)) == 0x8A9136AA
    assert T._crc32c_py(bytes([0xFF] * 32)) == 0x62A8AB43
    assert T._crc32c_py(bytes(range(32))) == 0x46DD794E
    big = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8)
    assert T.crc32c(big) == T._crc32c_py(big.tobytes())             # SSE4.2 routine of the library
    assert T.crc32c(big[5:], T.crc32c(big[:5])) == T.crc32c(big)      # continuation
    # leveldb's mask: rotate right by 15, add a constant
    assert T.mask_crc(0) == 0xA282EAD8


def test_roundtrip_many_blocks_and_prefix_compression(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {f"unet/_input_blocks/{i}/_residual/_conv1/kernel{T.SUFFIX}": rng.standard_normal((3, 3, 4, 5), dtype=np.float32)
               for i in range(40)}
    tensors["save_counter" + T.SUFFIX] = np.array(1, np.int64)
    tensors["unet/empty" + T.SUFFIX] = np.zeros((0, 7), np.float32)
    tensors["unet/scalar" + T.SUFFIX] = np.float32(3.5)
    prefix = str(tmp_path / "ckpt-1")
    T.write_checkpoint(prefix, tensors, block_size=512, restart_interval=4)   # many blocks, shared prefixes
    header, entries = T.read_index(prefix)
    assert header["num_shards"] == 1 and sorted(entries) == sorted(tensors)
    got = T.load_checkpoint(prefix)
    for k, v in tensors.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape
        assert np.array_equal(got[k], v)
    one = T.load_checkpoint(prefix, ["unet/scalar" + T.SUFFIX])
    assert list(one) == ["unet/scalar" + T.SUFFIX]
    with pytest.raises(T.CheckpointError):
        T.load_checkpoint(prefix, ["unet/missing" + T.SUFFIX])


def test_hand_assembled_index_is_parsed(tmp_path):
    """An index built byte by byte from the format description (independent of write_checkpoint)."""
    payload = np.arange(6, dtype=np.float32).reshape(2, 3)
    raw = payload.tobytes()
    entry = (b"\x08\x01" + b"\x12\x08" + b"\x12\x02\x08\x02" + b"\x12\x02\x08\x03" + b"\x28" + bytes([len(raw)]) +
             b"\x35" + struct.pack("<I", T.mask_crc(T._crc32c_py(raw))))
    header = b"\x08\x01\x1a\x02\x08\x01"
    k1, k2 = b"a/kernel", b"a/kernel2"           # k2 shares 8 bytes with k1
    block = (b"\x00\x00" + bytes([len(header)]) + header +
             b"\x00" + bytes([len(k1), len(entry)]) + k1 + entry +
             bytes([8, 1, len(entry)]) + b"2" + entry +
             struct.pack("<II", 0, 1))

    def with_trailer(b):
        return b + b"\x00" + struct.pack("<I", T.mask_crc(T._crc32c_py(b + b"\x00")))
    out = with_trailer(block)
    meta_off = len(out)
    meta = struct.pack("<II", 0, 1)
    out += with_trailer(meta)
    idx_off = len(out)
    handle = bytes([0, len(block)])
    idx = b"\x00" + bytes([len(k2), len(handle)]) + k2 + handle + struct.pack("<II", 0, 1)
    out += with_trailer(idx)
    footer = bytes([meta_off, len(meta), idx_off, len(idx)])
    out += footer + bytes(40 - len(footer)) + struct.pack("<Q", T.MAGIC)
    prefix = str(tmp_path / "hand")
    open(prefix + ".index", "wb").write(out)
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    got = T.load_checkpoint(prefix)
    assert sorted(got) == ["a/kernel", "a/kernel2"]
    assert np.array_equal(got["a/kernel2"], payload)


class _TfStyleTable:
    """A second, independent table writer used only by the test below: written from the LevelDB / TensorFlow
    table format description the way TensorFlow's own BundleWriter drives it (tensorflow/core/lib/io/
    table_builder.cc, block_builder.cc): prefix compression with a restart point every 16 entries, data blocks cut
    when they reach `block_size`, index keys shortened to the shortest separator between two blocks
    (BytewiseComparator::FindShortestSeparator) and the last one to a short successor, a metaindex block, a
    48-byte footer.  It shares no code with ldm_tf2_b200.tf_checkpoint."""

    def __init__(self, block_size):
        self.block_size, self.out, self.index = block_size, bytearray(), []
        self.entries, self.restarts, self.buf, self.last = 0, [0], bytearray(), b""
        self.pending = None   # (last key of the finished block, handle)

    @staticmethod
    def _v(n):
        o = bytearray()
        while n >= 0x80:
            o.append((n & 0x7F) | 0x80)
            n >>= 7
        o.append(n)
        return bytes(o)

    @staticmethod
    def _separator(a, b):
        n = 0
        while n < min(len(a), len(b)) and a[n] == b[n]:
            n += 1
        if n < min(len(a), len(b)) and a[n] < 0xFF and a[n] + 1 < b[n]:
            return a[:n] + bytes([a[n] + 1])
        return a

    def _crc(self, data):
        c = T._crc32c_py(bytes(data))
        return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF

    def _emit(self, block):
        off = len(self.out)
        self.out += block + b"\x00" + struct.pack("<I", self._crc(block + b"\x00"))
        return self._v(off) + self._v(len(block))

    def _finish_block(self):
        blk = bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))
        self.pending = (self.last, self._emit(blk))
        self.entries, self.restarts, self.buf = 0, [0], bytearray()

    def add(self, key, value):
        if self.pending:
            self.index.append((self._separator(self.pending[0], key), self.pending[1]))
            self.pending = None
        shared = 0
        if self.entries % 16 == 0 and self.entries:
            self.restarts.append(len(self.buf))
        elif self.entries:
            while shared < min(len(self.last), len(key)) and self.last[shared] == key[shared]:
                shared += 1
        self.buf += self._v(shared) + self._v(len(key) - shared) + self._v(len(value)) + key[shared:] + value
        self.last = key
        self.entries += 1
        if len(self.buf) >= self.block_size:
            self._finish_block()

    def finish(self):
        if self.entries:
            self._finish_block()
        if self.pending:
            k = self.pending[0]
            succ = next((k[:i] + bytes([k[i] + 1]) for i in range(len(k)) if k[i] != 0xFF), k)   # FindShortSuccessor
            self.index.append((succ, self.pending[1]))
        meta = self._emit(struct.pack("<II", 0, 1))
        ib = bytearray()
        restarts = []
        for k, h in self.index:   # index block: restart interval 1
            restarts.append(len(ib))
            ib += self._v(0) + self._v(len(k)) + self._v(len(h)) + k + h
        idx = self._emit(bytes(ib) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts)))
        foot = meta + idx
        return bytes(self.out) + foot + bytes(40 - len(foot)) + struct.pack("<Q", 0xDB4775248B80FB57)


def test_tensorflow_style_object_checkpoint_is_read(tmp_path):
    """What `tf.train.Checkpoint(unet=unet).save(prefix)` puts on disk (convert_ckpt_pytorch_to_tf2.py:426-431), hand
    assembled by the independent writer above: the bundle header, the `_CHECKPOINTABLE_OBJECT_GRAPH` string entry, the
    int64 `save_counter`, and the variables under their attribute paths in bytewise key order -- spread over many
    small data blocks (restart points inside blocks, shortened index keys).  The reader must find every variable
    the library asks for, skip the string entry, and verify every checksum."""
    from ldm_tf2_b200 import lib
    from oracle import ldm_oracle as O
    cfg = O.TINY_CONFIG
    d = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), -1)
    keys = T.variable_keys(d, d.UNET)
    shapes = [d.weight_info(d.UNET, i)[1] for i in range(d.num_weights(d.UNET))]
    rng = np.random.default_rng(0)
    tensors = {k: rng.standard_normal(s).astype(np.float32) for k, s in zip(keys, shapes)}
    graph = T.build_object_graph([k[:-len(T.SUFFIX)] for k in keys] + ["save_counter"])   # a DT_STRING scalar
    # ---- data file, in key order; a DT_STRING tensor is [varint lengths][masked crc of the lengths][bytes]
    allkeys = sorted([k.encode() for k in tensors] + [b"_CHECKPOINTABLE_OBJECT_GRAPH", b"save_counter" + T.SUFFIX.encode()])
    data = bytearray()
    table = _TfStyleTable(block_size=700)
    v = _TfStyleTable._v
    table.add(b"", b"\x08\x01\x1a\x02\x08\x01")   # BundleHeaderProto{num_shards 1, version{producer 1}}

    def crc_of(raw):
        c = T._crc32c_py(bytes(raw))
        return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF

    def shape_pb(shape):
        dims = b"".join(b"\x12" + v(len(dpb)) + dpb for dpb in (b"\x08" + v(int(n)) for n in shape))
        return b"\x12" + v(len(dims)) + dims

    for k in allkeys:
        off = len(data)
        if k == b"_CHECKPOINTABLE_OBJECT_GRAPH":
            # WriteStringTensor: the checksums run over the length as a uint32, not over its varint bytes
            lens = v(len(graph))
            c_len = T._crc32c_py(struct.pack("<I", len(graph)))
            cks = struct.pack("<I", (((c_len >> 15) | (c_len << 17)) + 0xA282EAD8) & 0xFFFFFFFF)
            raw = lens + cks + graph
            c_all = T._crc32c_py(graph, T._crc32c_py(cks, c_len))
            string_crc = (((c_all >> 15) | (c_all << 17)) + 0xA282EAD8) & 0xFFFFFFFF
            dtype, shp = 7, ()
        elif k.startswith(b"save_counter"):
            raw = np.int64(1).tobytes()
            dtype, shp = 9, ()
        else:
            a = tensors[k.decode()]
            raw = a.tobytes()
            dtype, shp = 1, a.shape
        data += raw
        entry = b"\x08" + v(dtype) + shape_pb(shp) + (b"\x20" + v(off) if off else b"") + b"\x28" + v(len(raw)) + \
            b"\x35" + struct.pack("<I", string_crc if dtype == 7 else crc_of(raw))
        table.add(k, entry)
    prefix = str(tmp_path / "unet-1")
    open(prefix + ".index", "wb").write(table.finish())
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    assert len(table.index) > 20                     # many data blocks, shortened separator keys in the index
    header, entries = T.read_index(prefix)
    assert header["num_shards"] == 1 and len(entries) == len(allkeys) - 0
    assert entries["_CHECKPOINTABLE_OBJECT_GRAPH"]["dtype"] == T.DT_STRING
    got = T.load_checkpoint(prefix, keys)           # exactly what restore() asks for
    for k in keys:
        assert np.array_equal(got[k], tensors[k]), k
    assert T.resolve_variable_keys(d, d.UNET, prefix) == keys   # through the object graph entry, checksums verified
    d.close()
    everything = T.load_checkpoint(prefix)          # the string entry is skipped, save_counter is read
    assert "_CHECKPOINTABLE_OBJECT_GRAPH" not in everything
    assert everything["save_counter" + T.SUFFIX] == 1 and everything["save_counter" + T.SUFFIX].dtype == np.int64
    # a flipped payload byte in the middle of the data file is caught by the entry checksum
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(T.CheckpointError, match="checksum"):
        T.load_checkpoint(prefix, keys)


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c")
    T.write_checkpoint(prefix, {"w" + T.SUFFIX: np.ones((64, 64), np.float32)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(T.CheckpointError, match="payload checksum"):
        T.load_checkpoint(prefix)
    assert T.load_checkpoint(prefix, verify=False)["w" + T.SUFFIX].shape == (64, 64)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(T.CheckpointError):
        T.read_index(prefix)
    open(prefix + ".index", "wb").write(b"not a table")
    with pytest.raises(T.CheckpointError):
        T.read_index(prefix)


@pytest.mark.parametrize("kind", ["kl", "vq"])
def test_variable_keys_match_the_reference_objects(kind):
    """The library's weight names (describe-only handle: no GPU needed) are the attribute paths of
    the reference's own layers, in flat Keras order."""
    gold = json.load(open(os.path.join(GOLD, "ckpt_keys_small.json")))
    cfg = lib.make_config(SMALL["cond_stage_model"], SMALL["unet"], SMALL["autoencoder_" + kind], kind, 8)
    h = lib.Handle(cfg, -1)
    try:
        for model, name in ((h.TEXT, "transformer"), (h.UNET, "unet"), (h.AE, "autoencoder_" + kind)):
            assert T.variable_keys(h, model) == gold[name]
        with pytest.raises(lib.LdmError, match="describe-only"):
            h.finalize()
    finally:
        h.close()


def test_save_and_reload_flat_weight_list(tmp_path):
    """save() under the reference's keys, then the keyed read restore() performs (shape-checked)."""
    cfg = lib.make_config(SMALL["cond_stage_model"], SMALL["unet"], SMALL["autoencoder_kl"], "kl", 8)
    h = lib.Handle(cfg, -1)
    try:
        rng = np.random.default_rng(2)
        shapes = [h.weight_info(h.AE, i)[1] for i in range(h.num_weights(h.AE))]
        weights = [rng.standard_normal(s, dtype=np.float32) for s in shapes]
        prefix = str(tmp_path / "autoencoder-1")
        T.save(h, weights, prefix, model=h.AE)
        keys = T.variable_keys(h, h.AE)
        got = T.load_checkpoint(prefix, keys)
        assert all(np.array_equal(got[k], w) for k, w in zip(keys, weights))
        assert keys[0] == "autoencoder/_post_quant_conv/kernel" + T.SUFFIX
    finally:
        h.close()


# ---------------------------------------------------------------------------------------------------------------
# Pins against TensorFlow's own code as shipped inside `tensorboard.compat` (the image has tensorboard, not TF):
# the masked CRC-32C routine the TF team wrote for record files, and the protobuf classes compiled from TF's
# trackable_object_graph.proto / tensor_shape.proto / types.proto.
# ---------------------------------------------------------------------------------------------------------------
def test_masked_crc_matches_tensorflows_own_routine():
    tb = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 8, 9, 63, 4096, 5000, 70001):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(data) == tb.crc32c(data) & 0xFFFFFFFF, n          # table path and SSE4.2 path of the library
        assert T.mask_crc(T.crc32c(data)) == tb.masked_crc32c(data) & 0xFFFFFFFF, n


def test_object_graph_parser_against_tensorflows_compiled_protos():
    pb = pytest.importorskip("tensorboard.compat.proto.trackable_object_graph_pb2")
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    # dtype enum values of types.proto
    for name in ("DT_FLOAT", "DT_DOUBLE", "DT_INT32", "DT_UINT8", "DT_STRING", "DT_INT64", "DT_BOOL", "DT_BFLOAT16", "DT_HALF"):
        assert getattr(T, name) == getattr(types_pb2, name), name
    # a graph written by TF's proto classes, with the fields the reader must ignore (slot variables of an optimizer,
    # full_name, has_checkpoint_values, registered_saver) and children listed out of node order
    g = pb.TrackableObjectGraph()
    root, unet, conv, kern, bias, ctr, opt = (g.nodes.add() for _ in range(7))
    root.children.add(node_id=6, local_name="optimizer")
    root.children.add(node_id=1, local_name="unet")
    root.children.add(node_id=5, local_name="save_counter")
    unet.children.add(node_id=2, local_name="_conv_in")
    conv.children.add(node_id=3, local_name="kernel")
    conv.children.add(node_id=4, local_name="bias")
    kern.attributes.add(name="VARIABLE_VALUE", full_name="u_net/conv2d/kernel",
                        checkpoint_key="unet/_conv_in/kernel/.ATTRIBUTES/VARIABLE_VALUE")
    bias.attributes.add(name="VARIABLE_VALUE", full_name="u_net/conv2d/bias",
                        checkpoint_key="unet/_conv_in/bias/.ATTRIBUTES/VARIABLE_VALUE")
    ctr.attributes.add(name="VARIABLE_VALUE", full_name="save_counter",
                       checkpoint_key="save_counter/.ATTRIBUTES/VARIABLE_VALUE")
    opt.slot_variables.add(original_variable_node_id=3, slot_name="m", slot_variable_node_id=3)
    opt.registered_saver.name = "none"
    kern.has_checkpoint_values.value = True
    nodes = T.parse_object_graph(g.SerializeToString())
    assert len(nodes) == 7
    assert nodes[0]["children"] == {"optimizer": 6, "unet": 1, "save_counter": 5}
    assert nodes[2]["children"] == {"kernel": 3, "bias": 4}
    assert T.resolve_key(nodes, "unet/_conv_in/kernel") == "unet/_conv_in/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert T.resolve_key(nodes, "save_counter") == "save_counter/.ATTRIBUTES/VARIABLE_VALUE"
    with pytest.raises(T.CheckpointError, match="no edge '_conv_out'"):
        T.resolve_key(nodes, "unet/_conv_out/kernel")
    with pytest.raises(T.CheckpointError, match="not a variable"):
        T.resolve_key(nodes, "unet/_conv_in")
    # an edge that points outside the node list is refused
    root.children.add(node_id=99, local_name="dangling")
    with pytest.raises(T.CheckpointError, match="points at node 99"):
        T.parse_object_graph(g.SerializeToString())
    # the other direction: this module's graph builder is byte-identical to TF's serializer for the same message
    paths = ["unet/_conv_in/kernel", "unet/_conv_in/bias", "unet/_blocks/0/_dense/kernel", "save_counter"]
    mine = T.build_object_graph(paths)
    m = pb.TrackableObjectGraph()
    m.ParseFromString(mine)
    assert m.SerializeToString() == mine and len(m.nodes) == 10
    assert [(c.local_name, c.node_id) for c in m.nodes[0].children] == [("unet", 1), ("save_counter", 2)]
    assert sorted(a.checkpoint_key for n in m.nodes for a in n.attributes) == sorted(p + T.SUFFIX for p in paths)
    # BundleEntryProto carries a TensorShapeProto: TF's serializer for the nested message, this module's entry parser
    shp = shape_pb2.TensorShapeProto()
    for d in (3, 3, 320, 1280):
        shp.dim.add(size=d)
    body = shp.SerializeToString()
    entry = b"\x08\x01" + b"\x12" + bytes([len(body)]) + body + b"\x20\x80\x01" + b"\x28\x10" + b"\x35" + struct.pack("<I", 7)
    e = T._parse_entry(entry)
    assert e == dict(dtype=T.DT_FLOAT, shape=(3, 3, 320, 1280), shard=0, offset=128, size=16, crc=7, sliced=False)
    mine_shape = T._build_entry(T.DT_FLOAT, (3, 3, 320, 1280), 128, 16, 7)
    assert mine_shape == entry


def test_restore_matches_objects_through_the_saved_graph(tmp_path):
    """Checkpoint.restore matches objects by walking the saved graph's edges, not by comparing key strings.  Here
    the saving program reached the UNet's first ResBlock over an alias attribute first (`unet/_alias`), so TensorFlow
    recorded those variables under `unet/_alias/...` keys; the loader asks for them under its own attribute paths
    and must get the right tensors.  The bundle is the TF-style object checkpoint (graph + save_counter)."""
    from oracle import ldm_oracle as O
    cfg = O.TINY_CONFIG
    d = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), -1)
    try:
        paths = [d.weight_info(d.UNET, i)[0] for i in range(d.num_weights(d.UNET))]
        shapes = [d.weight_info(d.UNET, i)[1] for i in range(d.num_weights(d.UNET))]
        block = "/".join(next(p for p in paths if p.count("/") >= 3).split("/")[:3])   # e.g. unet/_down_blocks/0
        aliased = [p for p in paths if p.startswith(block + "/")]
        assert aliased and len(aliased) < len(paths)

        def key_of(path):
            if path.startswith(block + "/"):
                return "unet/_alias" + path[len(block):] + T.SUFFIX
            return path + T.SUFFIX

        rng = np.random.default_rng(5)
        tensors = {key_of(p): rng.standard_normal(s).astype(np.float32) for p, s in zip(paths, shapes)}
        tensors["save_counter" + T.SUFFIX] = np.asarray(1, np.int64)
        graph = T.build_object_graph(paths + ["save_counter"], extra_edges=[("unet", "_alias", block)], key_of=key_of)
        prefix = str(tmp_path / "unet-1")
        T.write_checkpoint(prefix, tensors, block_size=900, strings={T.OBJECT_GRAPH_KEY: graph})
        nodes = T.read_object_graph(prefix)
        assert nodes[1]["children"]["_alias"] == nodes[nodes[1]["children"][block.split("/")[1]]]["children"][block.split("/")[2]]
        keys = T.resolve_variable_keys(d, d.UNET, prefix)
        assert keys == [key_of(p) for p in paths] and keys != T.variable_keys(d, d.UNET)
        got = T.load_checkpoint(prefix, keys)
        assert all(np.array_equal(got[k], tensors[k]) for k in keys)
        # name-based bundle (no graph entry): the literal attribute-path keys
        T.save(d, [tensors[key_of(p)] for p in paths], str(tmp_path / "named-1"), model=d.UNET)
        assert T.read_object_graph(str(tmp_path / "named-1")) is None
        assert T.resolve_variable_keys(d, d.UNET, str(tmp_path / "named-1")) == T.variable_keys(d, d.UNET)
        # save(object_graph=True) writes what Checkpoint.save writes; it resolves to the same keys
        T.save(d, [tensors[key_of(p)] for p in paths], str(tmp_path / "obj-1"), model=d.UNET, object_graph=True)
        assert T.resolve_variable_keys(d, d.UNET, str(tmp_path / "obj-1")) == T.variable_keys(d, d.UNET)
        assert T.load_checkpoint(str(tmp_path / "obj-1"))["save_counter" + T.SUFFIX] == 1
        # a corrupted graph payload: the string checksums catch it; resolve falls back to the literal keys with a warning
        data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
        off = T.read_index(prefix)[1][T.OBJECT_GRAPH_KEY]["offset"]
        data[off + 40] ^= 0x01
        open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
        with pytest.raises(T.CheckpointError, match="checksum"):
            T.read_object_graph(prefix)
        with pytest.warns(UserWarning, match="object graph unreadable"):
            assert T.resolve_variable_keys(d, d.UNET, prefix) == T.variable_keys(d, d.UNET)
    finally:
        d.close()


def test_string_tensor_layout_and_checksums():
    """tensor_bundle.cc WriteStringTensor: [varint64 lengths][masked CRC-32C of the lengths as uint32s][bytes]; the
    entry checksum runs over the uint32 lengths, the 4 checksum bytes and the string bytes (not over the varints)."""
    strings = [b"", b"a", b"x" * 300]
    raw, crc = T._string_payload(strings)
    assert raw[:4] == b"\x00\x01\xac\x02"                       # 0, 1, 300 as varints
    lens_crc = T._crc32c_py(struct.pack("<III", 0, 1, 300))
    assert struct.unpack_from("<I", raw, 4)[0] == T.mask_crc(lens_crc)
    running = T._crc32c_py(b"a" + b"x" * 300, T._crc32c_py(raw[4:8], lens_crc))
    assert crc == T.mask_crc(running)
    assert T._string_tensor(raw, 3, crc, True, "t") == strings
    with pytest.raises(T.CheckpointError, match="layout"):
        T._string_tensor(raw[:-1], 3, crc, True, "t")
    bad = bytearray(raw)
    bad[5] ^= 1
    with pytest.raises(T.CheckpointError, match="length checksum"):
        T._string_tensor(bytes(bad), 3, crc, True, "t")
    assert T._string_tensor(bytes(bad), 3, crc, False, "t") == strings
