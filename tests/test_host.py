"""CPU: host-side logic of the drop-in (no GPU compute): C ABI exports, error behaviour without
a device, schedule tables, DLPack unwrapping, reference-compatible signatures, sample sharding and
the world-size-2 image all-gather over gloo."""
import ast
import ctypes
import os
import re
import sys
import time

import numpy as np
import pytest

from oracle import ldm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from ldm_tf2_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(libpath):
    hdr = open(os.path.join(ROOT, "include", "ldm_b200.h")).read()
    declared = sorted(set(re.findall(r"LDM_API\s+[\w\s\*]+?\b(ldm_\w+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(libpath)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from ldm_tf2_b200 import lib as L
    assert sorted(L.EXPORTS) == declared  # the ctypes prototypes cover the whole header


def test_no_cpu_fallback_without_a_device(libpath):
    from ldm_tf2_b200 import lib as L
    cfg = O.TINY_CONFIG
    c = L.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8)
    try:
        h = L.Handle(c, 0)
    except L.LdmError as e:
        assert "-2" in str(e) and ("CUDA" in str(e) or "device" in str(e))
        return
    h.close()
    pytest.skip("a CUDA device is present: the no-device error path cannot be exercised here")


def test_bad_config_is_rejected(libpath):
    from ldm_tf2_b200 import lib as L
    cfg = {k: dict(v) for k, v in O.TINY_CONFIG.items()}
    cfg["unet"]["model_channels"] = 48  # not a multiple of 32
    c = L.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8)
    with pytest.raises(L.LdmError):
        L.Handle(c, 0)


@pytest.mark.parametrize("eta,S", [(0.0, 50), (1.0, 200), (0.3, 7)])
def test_host_schedule_equals_oracle_bitwise(eta, S):
    from ldm_tf2_b200.schedule import DDIMSchedule
    s = DDIMSchedule(1000, 0.00085, 0.012, 0.0, eta, S)
    o = O.ddim_schedule(eta=eta, num_ddim_steps=S, num_steps=1000, beta_start=0.00085, beta_end=0.012)
    assert np.array_equal(s.ddim_steps, o["ddim_steps"])
    table = s.coeff_table()
    for i in range(S):
        assert np.array_equal(table[i, :5], np.array(O.ddim_coeffs(o, i), np.float32))


def test_dlpack_borrow_host_tensors():
    import torch
    from ldm_tf2_b200.dlpack import borrow
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    x, shape, _ = borrow(a, np.float32)
    assert x is not None and tuple(shape) == (2, 3, 4) and np.array_equal(x, a)
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    x, shape, _ = borrow(t, np.float32)
    assert isinstance(x, np.ndarray) and np.array_equal(x, a)
    cap = torch.utils.dlpack.to_dlpack(torch.arange(6, dtype=torch.int64).reshape(2, 3))
    x, shape, _ = borrow(cap, np.int64)
    assert np.array_equal(x, np.arange(6).reshape(2, 3))
    with pytest.raises(ValueError):
        borrow(torch.utils.dlpack.to_dlpack(torch.zeros(2, dtype=torch.float64)), np.float32)


def _ref_args(path, cls, fn):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == fn:
                    return [a.arg for a in f.args.args if a.arg != "self"]
    raise KeyError((cls, fn))


def test_signatures_mirror_the_reference():
    """Drop-in: every argument of the reference's constructors / sampling methods exists, in the
    same order, in ours (ours may append keyword-only extras such as device, x_init, noise)."""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "model_runners.py")):
        pytest.skip("reference sources not present on this machine")
    import inspect
    from ldm_tf2_b200 import sampler as S
    pairs = [
        ("model_runners.py", "LatentDiffusionModel", "__init__", S.LatentDiffusionModelSampler.__init__),
        ("model_runners.py", "LatentDiffusionModelSampler", "ddim_sample", S.LatentDiffusionModelSampler.ddim_sample),
        ("model_runners.py", "LatentDiffusionModelSampler", "ddim_p_sample_loop", S.LatentDiffusionModelSampler.ddim_p_sample_loop),
        ("model_runners.py", "LatentDiffusionModelSampler", "ddim_p_sample_loop_progressive", S.LatentDiffusionModelSampler.ddim_p_sample_loop_progressive),
        ("model_runners.py", "LatentDiffusionModel", "decode_first_stage", S.LatentDiffusionModelSampler.decode_first_stage),
        ("unet.py", "UNet", "__init__", S.UNet.__init__),
        ("transformer.py", "TransformerModel", "__init__", S.TransformerModel.__init__),
        ("autoencoder.py", "AutoencoderKL", "__init__", S.AutoencoderKL.__init__),
        ("autoencoder.py", "AutoencoderVQ", "__init__", S.AutoencoderVQ.__init__),
    ]
    for f, cls, fn, ours in pairs:
        want = _ref_args(os.path.join(ref, f), cls, fn)
        have = [p for p in inspect.signature(ours).parameters if p != "self"]
        assert have[: len(want)] == want, (cls, fn, want, have)


def test_default_token_ids_layout():
    from ldm_tf2_b200 import tokens
    ids = tokens.default_token_ids(3)
    assert ids.shape == (6, 77) and ids.dtype == np.int64
    assert (ids[:3] == np.array(tokens.UNCOND_IDS)).all() and (ids[3:] == np.array(tokens.COND_IDS)).all()


def test_shard_ranges_cover_the_batch():
    from ldm_tf2_b200.parallel import shard_batch, shard_range, shard_token_ids
    for total in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    x = np.arange(8 * 2).reshape(8, 2)
    assert np.array_equal(np.concatenate([shard_batch(x, r, 4) for r in range(4)]), x)
    ids = np.arange(16 * 3).reshape(16, 3)  # 8 uncond rows then 8 cond rows
    s1 = shard_token_ids(ids, 1, 4)
    assert np.array_equal(s1, np.concatenate([ids[2:4], ids[10:12]]))


def _gloo_worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist
    from ldm_tf2_b200.parallel import allgather_images, shard_batch
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = np.random.default_rng(1234).standard_normal((total, 4, 4, 3)).astype(np.float32)  # global, seeded
    local = torch.from_numpy(shard_batch(g, rank, world) * 2.0 + 1.0)  # stand-in for sample + decode
    out = allgather_images(local, total)
    q.put((rank, np.array_equal(out.numpy(), g * 2.0 + 1.0)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_world_size_2_allgather_matches_single_process(total):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_wordpiece_matches_hf_golden(tmp_path):
    """run_ldm_sampler.py:28-46 without `transformers`: ids of 15 prompts (accents, CJK, emoji -> [UNK],
    punctuation, >77 tokens, a 120-character word) generated by HF's BertTokenizerFast on the reference's
    vocab.txt, including the two known-answer vectors of convert_ckpt_pytorch_to_tf2.py:384-392."""
    import json
    from ldm_tf2_b200 import tokens, wordpiece
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "wordpiece_small.json"), encoding="utf-8"))
    for prompt, ids in zip(g["prompts"], g["ids"]):
        assert wordpiece.encode(prompt, g["vocab"]) == ids, prompt[:40]
    assert wordpiece.encode(tokens.DEFAULT_PROMPT, g["vocab"]) == tokens.COND_IDS
    assert wordpiece.encode("", g["vocab"]) == tokens.UNCOND_IDS
    # vocab.txt on disk, the (uncond rows, cond rows) layout of get_token_ids
    inv = sorted(g["vocab"].items(), key=lambda kv: kv[1])
    lines = [""] * (inv[-1][1] + 1)
    for t, i in inv:
        lines[i] = t
    for i, t in enumerate(lines):
        if not t:
            lines[i] = f"[unused{i}]"
    (tmp_path / "vocab.txt").write_text("\n".join(lines) + "\n", encoding="utf-8")
    ids = tokens.get_token_ids(tokens.DEFAULT_PROMPT, str(tmp_path), 3)
    assert ids.dtype == np.int64 and ids.shape == (6, 77)
    assert (ids[:3] == np.array(tokens.UNCOND_IDS)).all() and (ids[3:] == np.array(tokens.COND_IDS)).all()


def test_bench_configs_and_reference_arm_plumbing(monkeypatch, capsys):
    """bench.py host logic without a GPU: the BASELINE.json configs shard over 1 / 2 / 4 / 8 ranks as strong
    scaling, the workload text names the BASELINE config, non-zero ranks of the reference arm exit without work, the
    BLAS pool is pinned to the cores the process may use (torchrun exports OMP_NUM_THREADS=1), and the checksum helper
    is the leading 16 hex digits of SHA-256."""
    import hashlib
    import importlib
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    assert set(bench.CONFIGS) == {"c2", "c3", "c4", "c5"} and bench.CONFIGS["c3"]["B"] == 64
    for world in (1, 2, 4, 8):
        assert bench.CONFIGS["c3"]["B"] % world == 0
        txt = bench.workload_text("c3", bench.CONFIGS["c3"], world, 64 // world, False)
        assert "configs[2]" in txt and "[64,32,32,4]" in txt and f"({64 // world}/GPU, strong scaling)" in txt
    assert "configs[1]" in bench.workload_text("c2", bench.CONFIGS["c2"], 1, 4, False)
    assert "eta=1" in bench.workload_text("c2", bench.CONFIGS["c2"], 1, 4, False)
    assert "512x512" in bench.workload_text("c4", bench.CONFIGS["c4"], 8, 1, False)
    assert "16384 codes" in bench.workload_text("c5", bench.CONFIGS["c5"], 1, 32, False)
    n = bench.cpu_threads()
    assert 1 <= n <= (os.cpu_count() or 1)
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    assert bench.sha16(a) == hashlib.sha256(a.tobytes()).hexdigest()[:16]
    # under torchrun only rank 0 runs the CPU arm: the other ranks return before touching the oracle
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2"])
    bench.main()
    assert capsys.readouterr().out == ""


def _free_port_block(n):
    """A base port with n consecutive free ports after it (the rendezvous probes base .. base + n - 1)."""
    import socket
    for _ in range(50):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        base = s.getsockname()[1]
        s.close()
        socks = []
        try:
            for p in range(base, base + n):
                t = socket.socket()
                t.bind(("127.0.0.1", p))
                socks.append(t)
            return base
        except OSError:
            continue
        finally:
            for t in socks:
                t.close()
    pytest.skip("no block of free ports")


@pytest.mark.parametrize("occupied", [0, 2])
def test_nccl_id_rendezvous_over_localhost(occupied):
    """parallel.exchange_bytes (the TCP rendezvous that carries the 128-byte NCCL id, no torch): world size 3 over
    localhost; with the first `occupied` ports of the list taken by a foreign listener that answers garbage, rank 0
    moves to the next port and the other ranks skip the foreign service (magic + job token in the hello)."""
    import socket
    import threading
    from ldm_tf2_b200 import parallel
    base = _free_port_block(parallel._PORT_TRIES)
    foreign, stop = [], threading.Event()

    def babble(srv):
        srv.settimeout(0.2)
        while not stop.is_set():
            try:
                c, _ = srv.accept()
            except (socket.timeout, OSError):
                continue
            with c:
                try:
                    c.sendall(b"HTTP/1.1 400 Bad Request\r\n\r\n")
                except OSError:
                    pass

    for p in range(base, base + occupied):
        s = socket.socket()
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", p))
        s.listen(4)
        foreign.append(s)
        threading.Thread(target=babble, args=(s,), daemon=True).start()
    payload = bytes(range(128))
    out = {}

    def run(rank):
        out[rank] = parallel.exchange_bytes(payload if rank == 0 else None, rank, 3, "127.0.0.1", base, timeout=30.0)

    ths = [threading.Thread(target=run, args=(r,)) for r in (2, 1, 0)]   # clients first: they retry until rank 0 listens
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=60)
    stop.set()
    for s in foreign:
        s.close()
    assert out == {0: payload, 1: payload, 2: payload}
    assert parallel.exchange_bytes(b"x", 0, 1) == b"x"   # world size 1: no socket at all


def test_rendezvous_rejects_another_jobs_token():
    """Two jobs with neighbouring base ports: a rank of job B that probes job A's port is not served A's id."""
    import socket
    import threading
    from ldm_tf2_b200 import parallel
    base = _free_port_block(parallel._PORT_TRIES + 1)
    got = {}

    def job_a(rank):
        got[("a", rank)] = parallel.exchange_bytes(b"A" * 128 if rank == 0 else None, rank, 2, "127.0.0.1", base, timeout=30.0)

    ta = threading.Thread(target=job_a, args=(0,))
    ta.start()
    time.sleep(0.3)   # A's rank 0 listens on `base`
    # a rank of job B (base port = base - 1 would probe `base` as its second candidate): same address, other token
    hello_b = parallel._MAGIC + parallel._job_token("127.0.0.1", base - 1, 2)
    with socket.create_connection(("127.0.0.1", base), timeout=5.0) as c:
        c.settimeout(5.0)
        c.sendall(hello_b)
        assert c.recv(16) == b""   # closed without a payload
    tb = threading.Thread(target=job_a, args=(1,))
    tb.start()
    ta.join(timeout=60)
    tb.join(timeout=60)
    assert got == {("a", 0): b"A" * 128, ("a", 1): b"A" * 128}


def test_reference_arm_line_has_the_contract_keys(monkeypatch, capsys):
    """bench.py --impl reference on rank 0: one JSON line with the same metric / unit / config as the CUDA arm,
    impl = reference, a cpu_baseline describing the run and a zero-copy e2e object; ms_per_step is the measured time of
    the bounded sample (so steps x ms_per_step fits the run), value the extrapolated metric.  The oracle is stubbed."""
    import importlib
    import json
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    calls = []

    def fake(name, steps, warmup):
        calls.append((name, steps, warmup))
        return dict(value=0.004, cores=8, step_ms=4500.0, sample="stub")

    monkeypatch.setattr(bench, "cpu_reference", fake)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip())
    assert calls == [("c3", 2, 1)]
    assert line["impl"] == "reference" and line["metric"] == bench.CONFIGS["c3"]["metric"] and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] == 0.004 and line["ms_per_step"] == 4500.0
    assert line["cpu_baseline"] == {"value": 0.004, "unit": "images/s", "cores": 8, "kind": "port", "sample": "stub"}
    assert line["e2e"] == {"value": 0.004, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[2]" in line["config"]["workload"] and "model" not in line["config"]
    # the reference arm runs "on the CUDA arm's config": the object is the one the CUDA arm prints for the same flags
    assert line["config"] == bench.config_dict("c3", bench.CONFIGS["c3"], 1, 64, False, 64)
    assert line["config"]["global_batch"] == 64 and line["config"]["baseline_config_index"] == 2


def test_no_undefined_names_in_the_host_code():
    """bench.py and the package's Python run on the GPU box where a typo only shows at round end: every name a
    function reads as a global must be a module-level name or a builtin (symtable walk, no execution)."""
    import builtins
    import symtable
    files = ["bench.py", "__graft_entry__.py"] + [os.path.join("ldm_tf2_b200", f) for f in sorted(os.listdir(os.path.join(ROOT, "ldm_tf2_b200"))) if f.endswith(".py")]
    files += [os.path.join("profiles", f) for f in sorted(os.listdir(os.path.join(ROOT, "profiles"))) if f.endswith(".py")]
    files += [os.path.join("tests", f) for f in sorted(os.listdir(os.path.join(ROOT, "tests"))) if f.endswith(".py")]
    for rel in files:
        src = open(os.path.join(ROOT, rel), encoding="utf-8").read()
        top = symtable.symtable(src, rel, "exec")
        known = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
        known |= {"__file__", "__name__", "__doc__"}
        bad = []

        def walk(t):
            for s in t.get_symbols():
                n = s.get_name()
                if s.is_global() and s.is_referenced() and n not in known and not hasattr(builtins, n):
                    bad.append((t.get_name(), n))
            for c in t.get_children():
                walk(c)

        walk(top)
        assert not bad, (rel, bad)


def test_shim_validates_buffer_shapes(libpath):
    """The C ABI takes raw pointers + a few sizes, so a buffer of the wrong shape would be read or written out of
    bounds: the Python shim refuses it before the call.  Runs on a describe-only handle (no GPU): a conforming call
    gets as far as the library and is refused there ("describe-only"), a non-conforming one never reaches it."""
    from ldm_tf2_b200 import lib
    cfg = O.TINY_CONFIG
    h = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], cfg["autoencoder_kl"], "kl", 8), -1)
    T, D = h.config.max_seq_len, h.config.context_dim
    z = lambda *s: np.zeros(s, np.float32)
    try:
        bad = [
            lambda: h.set_context(z(2, T - 1, D)),
            lambda: h.set_context(z(2, T, D + 8)),
            lambda: h.set_context(z(T, D)),
            lambda: h.unet_forward(z(2, 8, 8, 3), np.zeros(2, np.int32)),
            lambda: h.unet_forward(z(2, 8, 8, 4), np.zeros(3, np.int32)),
            lambda: h.configure_sampler(np.arange(5), z(4, 8)),
            lambda: h.configure_sampler(np.arange(5), z(5, 5)),
            lambda: h.ddim_step(z(2, 8, 8, 4), z(2, 8, 8, 4), None, 0, 5.0),           # eps needs 2B rows
            lambda: h.ddim_step(z(2, 8, 8, 4), z(4, 8, 8, 4), z(1, 8, 8, 4), 0, 5.0),  # noise of another batch
            lambda: h.sample(z(2, 8, 8), None, 5.0),
            lambda: h.sample(z(2, 8, 8, 4), z(2, 8, 8, 4), 5.0),                        # noise without the step axis
            lambda: h.sample(z(2, 8, 8, 4), None, 5.0, trace=True),                     # trace length unknown
            lambda: h.decode(z(2, 8, 8, 3)),
            lambda: h.encode_images(z(1, 64, 64, 4)),
            lambda: h.get_latents(z(1, 64, 64, 3), noise=z(1, 8, 8, 8)),
            lambda: h.vq_argmin(z(16, 3)),
            lambda: h.allgather(z(4), 8, z(8)),
            lambda: h.encode_text(np.zeros((2, T + 1), np.int64)),
        ]
        for i, f in enumerate(bad):
            with pytest.raises(lib.LdmError) as e:
                f()
            assert "describe-only" not in str(e.value), (i, str(e.value))
        good = [
            lambda: h.set_context(z(2, T, D)),
            lambda: h.unet_forward(z(2, 8, 8, 4), np.zeros(2, np.int32)),
            lambda: h.configure_sampler(np.arange(5), z(5, 8)),
            lambda: h.ddim_step(z(2, 8, 8, 4), z(4, 8, 8, 4), z(2, 8, 8, 4), 0, 5.0),
            lambda: h.sample(z(2, 8, 8, 4), z(5, 2, 8, 8, 4), 5.0, trace=True, num_steps=5),
            lambda: h.decode(z(2, 8, 8, 4)),
            lambda: h.encode_images(z(1, 64, 64, 3)),
            lambda: h.get_latents(z(1, 64, 64, 3), noise=z(1, 8, 8, 4)),
            lambda: h.vq_argmin(z(16, 4)),
            lambda: h.encode_text(np.zeros((2, T), np.int64)),
        ]
        for i, f in enumerate(good):
            with pytest.raises(lib.LdmError, match="describe-only"):
                f()
    finally:
        h.close()


def test_header_is_plain_c_and_links_from_a_c_host(libpath, tmp_path):
    """include/ldm_b200.h is the drop-in boundary: it must compile as plain C (what cgo / JNI / ctypes-style bindings
    see) and a C host must be able to drive the library through it.  No GPU: the program uses a describe-only handle
    (device -1), walks the UNet's weight table and checks that a compute call is refused with LDM_ERR_INVALID."""
    import shutil
    import subprocess
    from ldm_tf2_b200 import lib
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "ldm_b200.h"
int main(void) {
  ldm_config c;
  memset(&c, 0, sizeof c);
  c.vocab_size = 30522; c.encoder_stack_size = 2; c.hidden_size = 128; c.text_num_heads = 8; c.size_per_head = 16;
  c.max_seq_len = 77; c.filter_size = 256;
  c.model_channels = 64; c.out_channels = 4; c.num_blocks = 2; c.num_channel_mult = 4;
  c.channel_mult[0] = 1; c.channel_mult[1] = 2; c.channel_mult[2] = 4; c.channel_mult[3] = 4;
  c.num_heads = 8; c.head_base = 8; c.context_dim = 128;
  c.ae_kind = 0; c.latent_channels = 4; c.ae_channels = 32; c.ae_num_blocks = 2; c.ae_num_multipliers = 4;
  c.ae_multipliers[0] = 1; c.ae_multipliers[1] = 2; c.ae_multipliers[2] = 4; c.ae_multipliers[3] = 4;
  c.vq_vocab_size = 512; c.ae_build_latent_hw = 8; c.precision = 1;
  ldm_handle* h = NULL;
  if (ldm_create(&c, -1, &h) != LDM_OK) { printf("create failed: %s\n", ldm_last_error()); return 1; }
  int n = 0;
  if (ldm_num_weights(h, 1, &n) != LDM_OK) return 2;
  const char* name = NULL; int nd = 0; int shape[4];
  if (ldm_weight_info(h, 1, 0, &name, &nd, shape) != LDM_OK) return 3;
  printf("sizeof=%zu version=%d unet_weights=%d first=%s ndim=%d shape=%d,%d,%d,%d\n", sizeof(ldm_config), ldm_version(),
         n, name, nd, shape[0], shape[1], shape[2], shape[3]);
  int rc = ldm_finalize_weights(h);
  printf("finalize rc=%d msg=%s\n", rc, ldm_last_error());
  c.latent_channels = 5;
  ldm_handle* h2 = NULL;
  rc = ldm_create(&c, -1, &h2);
  printf("bad config rc=%d msg=%s\n", rc, ldm_last_error());
  return ldm_destroy(h) == LDM_OK ? 0 : 4;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(libpath)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-l:libldm_b200.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.splitlines()
    assert f"sizeof={C_sizeof_config()}" in out[0] and "version=200" in out[0], out[0]
    assert "unet_weights=686" in out[0] and "first=unet/_conv_in/kernel ndim=4 shape=3,3,4,64" in out[0], out[0]
    assert out[1].startswith(f"finalize rc={-1} ") and "describe-only" in out[1]
    assert out[2].startswith("bad config rc=-1") and "latent_channels" in out[2]


def C_sizeof_config():
    import ctypes
    from ldm_tf2_b200 import lib
    return ctypes.sizeof(lib.LdmConfig)


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the package imports it (AST walk over every import statement),
    bench.py imports it only inside cpu_reference (the CPU arm) and __graft_entry__ only inside smoke()."""
    def imports(path):
        tree = ast.parse(open(path, encoding="utf-8").read())
        out = []
        for fn in ast.walk(tree):
            scope = fn.name if isinstance(fn, (ast.FunctionDef, ast.AsyncFunctionDef)) else None
            body = ast.walk(fn) if scope else []
            for node in body:
                if isinstance(node, ast.Import):
                    out += [(scope, a.name) for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    out.append((scope, node.module or ""))
        top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
        for node in top:
            if isinstance(node, ast.Import):
                out += [(None, a.name) for a in node.names]
            else:
                out.append((None, node.module or ""))
        return out

    pkg = os.path.join(ROOT, "ldm_tf2_b200")
    for f in sorted(os.listdir(pkg)):
        if f.endswith(".py"):
            bad = [m for _, m in imports(os.path.join(pkg, f)) if m.split(".")[0] in ("oracle", "torch", "triton", "tensorflow")]
            # parallel.allgather_images (the gloo variant the CPU tests drive) is the one sanctioned torch import
            if f == "parallel.py":
                bad = [m for m in bad if m.split(".")[0] != "torch"]
            assert not bad, (f, bad)
    where = {s for s, m in imports(os.path.join(ROOT, "bench.py")) if m.split(".")[0] == "oracle"}
    assert where == {"cpu_reference"}, where
    where = {s for s, m in imports(os.path.join(ROOT, "__graft_entry__.py")) if m.split(".")[0] == "oracle"}
    assert where == {"smoke"}, where


@pytest.mark.parametrize("total,world", [(8, 2), (5, 2), (7, 4), (3, 8)])
def test_nccl_gather_host_logic_with_ragged_shards(total, world):
    """parallel.allgather_images_nccl: shards of unequal size are padded to the largest one for the collective and the
    padding is cut out again, so every rank ends up with the images in global order.  The handle is a stand-in whose
    allgather does what ncclAllGather does (rank r's `count` floats land at offset r * count)."""
    from ldm_tf2_b200 import parallel
    g = np.random.default_rng(9).standard_normal((total, 4, 4, 3)).astype(np.float32)
    shards = [parallel.shard_batch(g, r, world) for r in range(world)]
    maxn = max(s.shape[0] for s in shards)
    per = 4 * 4 * 3

    class FakeHandle:
        def __init__(self, rank):
            self._world, self._rank = world, rank

        def allgather(self, send, count, out):
            assert count == maxn * per and send.shape[0] == maxn and out.size == world * count
            assert np.array_equal(send[: shards[self._rank].shape[0]], shards[self._rank])
            assert not send[shards[self._rank].shape[0]:].any()          # zero padding
            flat = out.reshape(world, maxn, 4, 4, 3)
            flat[:] = 0
            for r, s in enumerate(shards):
                flat[r, : s.shape[0]] = s

    for r in range(world):
        if shards[r].shape[0] == 0 and total < world:
            local = np.zeros((0, 4, 4, 3), np.float32)
        else:
            local = shards[r]
        out = parallel.allgather_images_nccl(FakeHandle(r), local, total)
        assert out.shape == g.shape and np.array_equal(out, g)
    assert parallel.gather_plan(total, world, per)[0] == maxn
