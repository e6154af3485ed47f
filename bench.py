#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: 256x256 images/s, 50-step DDIM + CFG).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference
    python bench.py --config c2|c4|c5                        # the other BASELINE.json configs

One "step" = one full sampling job of this rank's shard of the global batch: S DDIM steps (each = one
CFG-doubled UNet forward + the fused CFG/DDIM update) + decode (+ the all-gather of the images when
N > 1).  Workloads (SURVEY 8a legend; BASELINE.json `configs` index in brackets):

    c3 [2] (default)  latent [64,32,32,4], 50 steps eta 0, guidance 5, KL decode  -- the config the
                      metric is quoted on.  STRONG scaling: the 64 images shard over the N GPUs
                      (64/32/16/8 per GPU).  --weak B keeps B images per GPU instead.
    c2 [1]            latent [4,32,32,4], 200 steps eta 1 (per-step noise), guidance 5, KL decode
    c4 [3]            latent [8,64,64,4] (4096-token self-attention), 50 steps, KL decode to 512x512
    c5 [4]            decode only, batch 32 at 256x256: KL decode and VQ decode (codebook argmin);
                      `index_match` = fraction of VQ indices equal to the CPU oracle's

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic work (SURVEY 8d / BASELINE.md section 2), 2*MAC, unpadded dims, per image
GFLOP_UNET_STEP = {32: 364.15, 64: 1610.48}   # one CFG step (2 UNet passes) incl. the context K/V projections
GFLOP_CTX_KV = 9.84                           # ... of which loop-invariant (hoisted into set_context)
# the part of a CFG step the implicit-GEMM kernel executes (conv3x3 + dense; attention products excluded):
# conv 55.0 % + dense 37.6 % of 364.15 at 32x32 minus the hoisted 9.84 (the library's own count of
# 2*M*N*K over one step's launches gives the same: profiles/r1_gemm_shapes.txt, 2696.5 GFLOP for 8 images)
GFLOP_GEMM_STEP = {32: 337.06, 64: None}
GFLOP_KL_DECODE = {32: 622.19, 64: 2514.52}
GFLOP_VQ_DECODE = 472.53
SCALE_FACTOR = 0.18215

CONFIGS = {
    "c2": dict(index=1, B=4, hw=32, steps=200, eta=1.0, metric="images_per_s_256x256_ddim200_eta1_cfg"),
    "c3": dict(index=2, B=64, hw=32, steps=50, eta=0.0, metric="images_per_s_256x256_ddim50_cfg"),
    "c4": dict(index=3, B=8, hw=64, steps=50, eta=0.0, metric="images_per_s_512x512_ddim50_cfg"),
    "c5": dict(index=4, B=32, hw=32, steps=0, eta=0.0, metric="images_per_s_256x256_decode_only"),
}
UNIT = "images/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    tflops_burst=float(d.get("bf16_tflops", 1590.0)), hbm=float(d.get("hbm_gbs", 6650.0)),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def workload_text(name, c, world, per_gpu, weak):
    if name == "c5":
        return (f"BASELINE configs[4]: decode only, batch {c['B']} at 256x256: KL decode + VQ decode with the "
                f"quantize.py codebook argmin (16384 codes), txt2img-f8-large random-init autoencoders")
    out = 8 * c["hw"]
    return (f"BASELINE configs[{c['index']}]: txt2img-f8-large random-init, latent [{per_gpu * world if weak else c['B']},"
            f"{c['hw']},{c['hw']},4] ({per_gpu}/GPU, {'weak' if weak else 'strong'} scaling), {c['steps']} DDIM steps "
            f"eta={c['eta']:g} + CFG (guidance 5), KL decode to {out}x{out}"
            + (", all-gather of the images" if world > 1 else ""))


def config_dict(name, c, world, per_gpu, weak, global_b):
    """The `config` object of the JSON line: the WORKLOAD, identical for the CUDA arm and the reference arm run with
    the same flags (arm-specific facts -- operand format, how time was taken -- are top-level keys of the line)."""
    return {"workload": workload_text(name, c, world, per_gpu, weak), "baseline_config_index": c["index"],
            "global_batch": global_b, "per_gpu_batch": per_gpu,
            "parallelism": f"dp{world} (sample-sharded weight replicas, no collective inside the loop)",
            "l2": "not flushed explicitly: every UNet step streams the whole weight set (1.75 GB as 16-bit operands, "
                  "3.5 GB in fp32) plus activations far larger than the 126 MB L2"}


# --------------------------------------------------------------------------------------
# CPU arm: the NumPy restatement of the reference (oracle/), the only place bench.py runs it
# --------------------------------------------------------------------------------------
def cpu_threads():
    """Pins the BLAS pool to every core this process may use (torchrun exports OMP_NUM_THREADS=1) and
    returns the count actually in effect."""
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        got = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(got)) if got else want
    except Exception:
        return want


def cpu_reference(name, steps, warmup):
    """A bounded sample of the workload on the host cores with the oracle (kind = "port": TensorFlow
    cannot be installed here).  Sampling configs: `steps` CFG UNet steps at B = 1 + one decode,
    extrapolated to S steps + decode per image.  c5: KL + VQ decode of ONE image."""
    from oracle import ldm_oracle as O
    threads = cpu_threads()
    c = CONFIGS[name]
    cfg = O.FULL_CONFIG
    hw = c["hw"]
    as_ = O.ae_spec(cfg["autoencoder_kl"], "kl")
    Wa = O.as_dict(as_, O.init_weights(as_, 2))
    rng = np.random.default_rng(1234)
    if name == "c5":
        vs = O.ae_spec(cfg["autoencoder_vq"], "vq")
        Wv = O.as_dict(vs, O.init_weights(vs, 3))
        z = np.random.default_rng(6).standard_normal((c["B"], hw, hw, 4), dtype=np.float32)
        t0 = time.perf_counter()
        O.decode_first_stage(Wa, cfg["autoencoder_kl"], "kl", z[:1])
        t_kl = time.perf_counter() - t0
        t0 = time.perf_counter()
        _, idx = O.vq_lookup((z / np.float32(SCALE_FACTOR)).astype(np.float32), Wv["autoencoder/_quantize/kernel"])
        t_idx = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.decode_first_stage(Wv, cfg["autoencoder_vq"], "vq", z[:1])
        t_vq = time.perf_counter() - t0
        per_image = t_kl + t_vq   # one KL decode + one VQ decode (argmin of its rows included)
        return dict(value=1.0 / per_image, cores=threads, vq_indices=idx, step_ms=per_image * 1e3,
                    sample=f"1 of {c['B']} images: KL decode {t_kl:.2f} s + VQ decode incl. argmin {t_vq:.2f} s "
                           f"(argmin of all {idx.size} rows: {t_idx:.2f} s); NumPy fp32 + OpenBLAS, {threads} threads")
    us = O.unet_spec(cfg["unet"])
    Wu = O.as_dict(us, O.init_weights(us, 0))
    sched = O.ddim_schedule(**dict(cfg["ldm"], eta=c["eta"], num_ddim_steps=c["steps"]))
    S = c["steps"]
    xt = rng.standard_normal((1, hw, hw, 4), dtype=np.float32)
    ctx = np.random.default_rng(3).standard_normal((2, 77, 1280), dtype=np.float32)
    nz = rng.standard_normal((1, hw, hw, 4), dtype=np.float32) if c["eta"] > 0 else None

    def one_step(x, index):
        t = np.full([2], sched["ddim_steps"][index], np.int32)
        e = O.unet_forward(Wu, cfg["unet"], np.concatenate([x, x]), t, ctx)
        return O.ddim_update(x, e[:1], e[1:], nz, O.ddim_coeffs(sched, index), 5.0)[0]

    t0 = time.perf_counter()
    O.decode_first_stage(Wa, cfg["autoencoder_kl"], "kl", xt * np.float32(SCALE_FACTOR))
    t_dec = time.perf_counter() - t0
    for i in range(warmup):
        xt = one_step(xt, S - 1 - i % S)
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        xt = one_step(xt, S - 1 - (warmup + i) % S)
        ts.append(time.perf_counter() - t0)
    t_step = float(np.mean(ts))
    per_image = S * t_step + t_dec
    return dict(value=1.0 / per_image, cores=threads, step_ms=t_step * 1e3,
                sample=f"B=1 at {hw}x{hw} latents: {steps} CFG UNet steps ({t_step:.2f} s each) + 1 KL decode "
                       f"({t_dec:.2f} s), extrapolated to {S} steps + decode per image; NumPy fp32 + OpenBLAS, "
                       f"{threads} threads")


def sha16(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--weak", type=int, default=0, metavar="B",
                    help="weak scaling: B images per GPU (round-1 line: --weak 8) instead of sharding the config's batch")
    ap.add_argument("--verify", action="store_true",
                    help="multi-GPU invariance: rank 0 recomputes every shard locally and compares it bit for bit "
                         "with the gathered images")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rooflines", action="store_true", help="skip the microbenchmarks behind the roofline keys")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.config
    c = CONFIGS[name]
    weak = args.weak > 0
    if weak:
        per_gpu, global_b = args.weak, args.weak * world
    else:
        global_b = c["B"]
        if global_b % world:
            raise SystemExit(f"config {name}: global batch {global_b} does not shard over {world} GPUs")
        per_gpu = global_b // world
    workload = workload_text(name, c, world, per_gpu, weak)
    scaling = "weak" if weak else "strong"

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(name, max(args.steps, 1), min(args.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": c["metric"], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["step_ms"], "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(name, c, world, per_gpu, weak, global_b),
            "note": "CPU restatement of the reference (oracle/, NumPy fp32) on a bounded sample of this workload; "
                    "TensorFlow is not installable here, so kind=port, not the TF2 sampler.  A step of this arm = the "
                    "bounded sample (ms_per_step is its measured wall time); value = the workload's metric extrapolated "
                    "from it (cpu_baseline.sample says how)",
            "timing": "time.perf_counter around each bounded step on the host",
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # keep stdout clean for the single JSON line: libraries (NCCL banner, ...) print to fd 1
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from ldm_tf2_b200 import lib, parallel, synth, tokens
    from ldm_tf2_b200.sampler import (AutoencoderKL, AutoencoderVQ, LatentDiffusionModelSampler, TransformerModel,
                                      UNet)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.FULL_CONFIG
    precision = os.environ.get("LDM_B200_PRECISION", lib.DEFAULT_PRECISION)
    peaks = measured_peaks()
    B, hw, S = per_gpu, c["hw"], c["steps"]
    out_hw = 8 * hw

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_int(v):
        if world == 1:
            return int(v)
        t = torch.tensor([v], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        return int(t.item())

    ldm_kw = dict(cfg["ldm"], eta=c["eta"], num_ddim_steps=max(S, 1))
    text = TransformerModel(**cfg["cond_stage_model"])
    unet = UNet(**cfg["unet"])
    ae = AutoencoderKL(**cfg["autoencoder_kl"])
    # public API objects, exactly as run_ldm_sampler.py:56-83 builds them
    sampler = LatentDiffusionModelSampler(unet, ae, text, device=local_rank, **ldm_kw)
    h = sampler.handle
    models = ((h.AE, 2),) if name == "c5" else ((h.TEXT, 1), (h.UNET, 0), (h.AE, 2))
    for model, seed in models:
        h.set_weights(model, synth.random_weights(h, model, seed))
    h.finalize()
    hv = None
    if name == "c5" or (rank == 0 and not args.no_rooflines):
        # VQ autoencoder (decode side + codebook) on its own handle: K6 lives there
        vq = AutoencoderVQ(**cfg["autoencoder_vq"])
        hv = lib.Handle(lib.make_config(cfg["cond_stage_model"], cfg["unet"], vq.kwargs, "vq", 32), local_rank)
        hv.set_weights(hv.AE, synth.random_weights(hv, hv.AE, 3))
        hv.finalize()

    extra = {}
    gathered = None
    gather_buf = None
    if world > 1:
        # NCCL communicator owned by the library (comm.cu); the id travels over a TCP rendezvous, not torch
        parallel.init_comm(h, rank, world)
        gather_buf = torch.empty((global_b, out_hw, out_hw, 3), dtype=torch.float32, device=dev)

    def gather(img_dev):
        """The one collective of the path: ldm_allgather_images over NCCL on the library stream;
        returns (gathered tensor, device ms from CUDA events around the collective)."""
        if world == 1:
            return img_dev, 0.0
        h.allgather(lib.DevPtr(img_dev.data_ptr(), img_dev.shape), img_dev.numel(),
                    lib.DevPtr(gather_buf.data_ptr(), gather_buf.shape))
        return gather_buf, h.timing_ex("gather")

    if name == "c5":
        # ------------------------------------------------------------------ decode-only workload
        zg = np.random.default_rng(6).standard_normal((global_b, hw, hw, 4), dtype=np.float32)
        z_np = parallel.shard_batch(zg, rank, world)
        z_dev = torch.from_numpy(z_np).to(dev)
        img_kl = torch.empty((B, out_hw, out_hw, 3), dtype=torch.float32, device=dev)
        img_vq = torch.empty_like(img_kl)
        idx_dev = torch.empty((B * hw * hw,), dtype=torch.int64, device=dev)

        def device_step():
            lib.check(h.lib.ldm_decode(h._h, lib.ptr(z_dev.data_ptr()), B, hw, hw, SCALE_FACTOR,
                                       lib.ptr(img_kl.data_ptr()), None))
            t_kl = h.timing()["decode_ms"]
            lib.check(hv.lib.ldm_decode(hv._h, lib.ptr(z_dev.data_ptr()), B, hw, hw, SCALE_FACTOR,
                                        lib.ptr(img_vq.data_ptr()), lib.ptr(idx_dev.data_ptr())))
            t_vq = hv.timing()["decode_ms"]
            _, g_ms = gather(img_kl)
            return t_kl + t_vq + g_ms, dict(loop_ms=0.0, decode_ms=t_kl, vq_ms=t_vq)

        def e2e_step():
            a = sampler.decode_first_stage(z_np)
            b_, idx = hv.decode(z_np, div=SCALE_FACTOR)
            return a, b_, idx

        h2d, d2h = 2 * z_np.nbytes, 2 * B * out_hw * out_hw * 3 * 4 + B * hw * hw * 8
        gflop_per_image = GFLOP_KL_DECODE[hw] + GFLOP_VQ_DECODE
    else:
        # ------------------------------------------------------------------ sampling workloads
        h.configure_sampler(sampler.schedule.ddim_steps, sampler.schedule.coeff_table())
        ids_g = tokens.default_token_ids(global_b)
        ids = parallel.shard_token_ids(ids_g, rank, world)
        # globally seeded x_T / noise, sliced per rank: results do not depend on the GPU count
        xg = np.random.default_rng(1234).standard_normal((global_b, hw, hw, 4), dtype=np.float32)
        x_host = torch.empty((B, hw, hw, 4), dtype=torch.float32).pin_memory()
        x_host.copy_(torch.from_numpy(parallel.shard_batch(xg, rank, world)))
        x_np = x_host.numpy()
        x_dev = x_host.to(dev)
        nz_np, nz_dev = None, None
        if c["eta"] > 0:
            ng = np.random.default_rng(5678).standard_normal((S, global_b, hw, hw, 4), dtype=np.float32)
            nz_np = parallel.shard_batch(ng, rank, world, axis=1)
            nz_dev = torch.from_numpy(nz_np).to(dev)
        img_dev = torch.empty((B, out_hw, out_hw, 3), dtype=torch.float32, device=dev)
        ctx = h.encode_text(ids)
        h.set_context(ctx)

        def device_step():
            """inputs resident in HBM: loop + decode of the device-resident latents (+ all-gather); device ms"""
            lib.check(h.lib.ldm_sample(h._h, lib.ptr(x_dev.data_ptr()), lib.ptr(None if nz_dev is None else nz_dev.data_ptr()),
                                       B, hw, hw, 5.0, None, None, 0, 1))
            lib.check(h.lib.ldm_decode(h._h, None, B, hw, hw, SCALE_FACTOR, lib.ptr(img_dev.data_ptr()), None))
            t = h.timing()
            nonlocal gathered
            gathered, g_ms = gather(img_dev)
            return t["loop_ms"] + t["decode_ms"] + g_ms, t

        def e2e_step():
            # the call a user makes: host token ids + host x_T (+ host noise) in, host images out
            return sampler.ddim_p_sample_loop(ids, (B, hw, hw, 4), 5.0, x_init=x_np, noise=nz_np)

        h2d = x_np.nbytes + ids.nbytes + ctx.nbytes + (nz_np.nbytes if nz_np is not None else 0)
        d2h = B * out_hw * out_hw * 3 * 4 + ctx.nbytes   # images + the text context (returned, then re-uploaded)
        gflop_per_image = S * (GFLOP_UNET_STEP[hw] - GFLOP_CTX_KV) + GFLOP_KL_DECODE[hw]

    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = h.timing()["launches"] + (hv.timing()["launches"] if hv else 0)
    w0 = time.perf_counter()
    dev_ms, acc = 0.0, {}
    for _ in range(args.steps):
        ms, t = device_step()
        dev_ms += ms
        for k, v in t.items():
            if k.endswith("_ms"):
                acc[k] = acc.get(k, 0.0) + v
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    launches = h.timing()["launches"] + (hv.timing()["launches"] if hv else 0) - l0
    clk = clocks.stop()

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_out = e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - e0) * 1e3

    dev_ms, wall_ms, e2e_ms = allmax(dev_ms), allmax(wall_ms), allmax(e2e_ms)
    launches_total = allsum_int(launches)

    # ---- multi-GPU invariance: gathered == every shard recomputed on this GPU, bit for bit
    if name != "c5":
        img_all = gathered if world > 1 else img_dev
        extra["images_sha256_16"] = sha16(img_all.cpu().numpy())
        if args.verify and rank == 0 and world > 1:
            ok = True
            all_np = img_all.cpu().numpy()
            chk = torch.empty_like(img_dev)
            for r in range(world):
                ids_r = parallel.shard_token_ids(ids_g, r, world)
                h.set_context(h.encode_text(ids_r))
                xr = np.ascontiguousarray(parallel.shard_batch(xg, r, world))
                nr = None if c["eta"] == 0 else np.ascontiguousarray(parallel.shard_batch(ng, r, world, axis=1))
                lib.check(h.lib.ldm_sample(h._h, lib.ptr(xr), lib.ptr(nr), B, hw, hw, 5.0, None, None, 0, 1))
                lib.check(h.lib.ldm_decode(h._h, None, B, hw, hw, SCALE_FACTOR, lib.ptr(chk.data_ptr()), None))
                lo, hi = parallel.shard_range(global_b, r, world)
                ok = ok and np.array_equal(all_np[lo:hi].view(np.uint32), chk.cpu().numpy().view(np.uint32))
            extra["verify"] = {"gathered_equals_local_recompute_bitwise": bool(ok), "shards": world}
            h.set_context(ctx)

    if rank == 0:
        total_images = global_b * args.steps
        value = total_images / (dev_ms / 1e3)
        line = {
            "metric": c["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": precision, "data": "synthetic",
            "config": config_dict(name, c, world, per_gpu, weak, global_b),
            "operands": f"{precision} tensor-core operands, fp32 accumulate and statistics, "
                        + ("fp32" if os.environ.get("LDM_B200_STREAM") == "fp32" else "16-bit") + " residual stream between blocks",
            "timing": "CUDA events on the library stream (loop, decode, NCCL all-gather), max over ranks",
            "wall_ms_per_step": wall_ms / args.steps,
            "model_tflops_per_gpu": value * gflop_per_image / 1e3 / world,
            "e2e": {"value": total_images / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "api": ("LatentDiffusionModelSampler.decode_first_stage (KL) + Handle.decode (VQ), host latents in, host images out"
                            if name == "c5" else
                            "LatentDiffusionModelSampler.ddim_p_sample_loop(ids, shape, guidance) incl. text encoder, host x_T in, host images out")},
            "gpu_launches": int(launches_total),
            "clocks": clk,
        }
        line.update(extra)
        if name == "c5":
            line["ms_decode_kl"] = acc["decode_ms"] / args.steps
            line["ms_decode_vq"] = acc["vq_ms"] / args.steps
            line["images_per_s_kl"] = B / (line["ms_decode_kl"] / 1e3) * world
            line["images_per_s_vq"] = B / (line["ms_decode_vq"] / 1e3) * world
            kl_tf = B * GFLOP_KL_DECODE[hw] / line["ms_decode_kl"]
            line["roofline"] = {"bound": "tensor", "kernel": "implicit_gemm_kernel (tcgen05), KL decoder", "achieved": kl_tf,
                                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": kl_tf / peaks["tflops"], "traffic": None,
                                "peak_source": peaks["source"] + ", sustained bf16",
                                "algorithmic_gflop_per_image": GFLOP_KL_DECODE[hw],
                                "note": "whole-decode time (conv3x3 = 97.9 % of the decoder FLOPs)"}
        else:
            ms_step = acc["loop_ms"] / args.steps / S
            line["ms_per_unet_step"] = ms_step
            line["ms_decode"] = acc["decode_ms"] / args.steps
            step_gflop = B * (GFLOP_UNET_STEP[hw] - GFLOP_CTX_KV)
            # north_star's number: the whole CFG UNet step (GEMMs, attention, norms, K5) against dense-bf16 peak
            line["roofline_step"] = {"bound": "tensor", "what": "one CFG UNet step incl. attention, norms and the K5 update",
                                     "achieved": step_gflop / ms_step, "peak": peaks["tflops"], "unit": "TFLOP/s",
                                     "frac": step_gflop / ms_step / peaks["tflops"],
                                     "frac_of_burst": step_gflop / ms_step / peaks["tflops_burst"],
                                     "algorithmic_gflop_per_step": step_gflop, "peak_source": peaks["source"] + ", sustained bf16"}
        if not args.no_rooflines:
            rooflines(line, h, hv, B, hw, name, peaks)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(name, 2, 1)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
            if name == "c5":
                got = e2e_out[2]
                line["index_match"] = float(np.mean(got == r["vq_indices"]))
                line["index_rows"] = int(got.size)
        else:
            line["cpu_baseline"] = None
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if hv:
        hv.close()
    sampler.close()


def rooflines(line, h, hv, B, hw, name, peaks):
    """Live microbenchmarks behind the roofline keys (rank 0, after the timed region)."""
    if name != "c5":
        # Dominant kernel: implicit_gemm_kernel.  Its time inside the replayed step graph = full step -
        # the same graph captured without its launches (CUDA events around 20 replays each); numerator =
        # the GEMM-only algorithmic FLOPs (attention products are executed by flash_attention_kernel).
        prof = h.profile_unet_step(B, hw, hw, 2)
        t_full = h.bench_unet_step(B, hw, hw, 20, True)
        t_rest = h.bench_unet_step(B, hw, hw, 20, True, skip_gemm=True)
        gemm_ms = t_full - t_rest
        gemm_gflop = B * GFLOP_GEMM_STEP[hw] if GFLOP_GEMM_STEP.get(hw) else prof["gemm_flops_per_step"] / 1e9
        achieved = gemm_gflop / gemm_ms
        line["roofline"] = {
            "bound": "tensor", "kernel": "implicit_gemm_kernel (tcgen05)", "achieved": achieved, "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "frac_of_burst": achieved / peaks["tflops_burst"],
            "traffic": GEMM_DRAM_BYTES_PER_UNET_STEP.get((hw, B)),
            "traffic_unit": "DRAM bytes per UNet step summed over all GEMM launches (profiles/r2_launches_unet_step_b64_summary.txt); "
                            "null for shapes without an ncu capture",
            "peak_source": peaks["source"] + ", sustained bf16",
            "algorithmic_gflop_per_unet_step": gemm_gflop,
            "executed_gflop_per_unet_step": prof["gemm_flops_per_step"] / 1e9,
            "launches_per_unet_step": prof["gemm_launches_per_step"],
            "kernel_ms_per_unet_step": gemm_ms, "step_ms_graph": t_full, "step_ms_graph_without_gemm": t_rest,
            "kernel_share_of_step": gemm_ms / t_full,
            "kernel_ms_event_per_launch_sum": prof["gemm_ms_per_step"], "eager_step_ms": prof["step_ms"]}
        k5_ms = h.bench_ddim_update(B, hw, hw, False, 200)
        k5_bytes = 4 * 4 * B * hw * hw * 4  # 3 reads + 1 write of fp32 [B,h,w,4]
        line["roofline_k5"] = {"bound": "hbm", "kernel": "ddim_update_kernel", "achieved": k5_bytes / (k5_ms * 1e-3) / 1e9,
                               "peak": peaks["hbm"], "unit": "GB/s", "frac": k5_bytes / (k5_ms * 1e-3) / 1e9 / peaks["hbm"],
                               "traffic": None, "bytes_per_launch": k5_bytes, "ms_per_launch": k5_ms}
    # K2 GroupNorm at the decoder's largest activation [8, 256*256, 128], the kernels the sampling path runs: 16-bit
    # residual stream in (134 MB, several buffers so that it is HBM-resident), 16-bit MMA operand out
    gn_n, gn_hw, gn_c = 8, 256 * 256, 128
    gn_stats_ms, gn_apply_ms = h.bench_groupnorm(gn_n, gn_hw, gn_c, 10, in16=True)
    f32_stats_ms, f32_apply_ms = h.bench_groupnorm(gn_n, gn_hw, gn_c, 10, in16=False)
    gn_el = gn_n * gn_hw * gn_c

    def _gb(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9
    line["roofline_k2"] = {
        "bound": "hbm", "kernel": "gn_stats16x8_kernel + gn_apply16x8_kernel (GroupNorm(32)+SiLU, 16-bit stream -> 16-bit operand)",
        "shape": [gn_n, gn_hw, gn_c], "unit": "GB/s", "peak": peaks["hbm"],
        "stats": {"bytes_per_launch": gn_el * 2, "ms_per_launch": gn_stats_ms, "achieved": _gb(gn_el * 2, gn_stats_ms),
                  "frac": _gb(gn_el * 2, gn_stats_ms) / peaks["hbm"]},
        "apply": {"bytes_per_launch": gn_el * 4, "ms_per_launch": gn_apply_ms, "achieved": _gb(gn_el * 4, gn_apply_ms),
                  "frac": _gb(gn_el * 4, gn_apply_ms) / peaks["hbm"]},
        "fp32_input_flavour": {"stats_gbs": _gb(gn_el * 4, f32_stats_ms), "apply_gbs": _gb(gn_el * 6, f32_apply_ms),
                               "note": "LDM_B200_STREAM=fp32 / text-free paths: 4 B read (stats), 4 B read + 2 B write (apply)"},
        "traffic": K2_NCU_TRAFFIC}
    if hv is not None:
        # K6: rows x 16384 codes, exact fp32 op order without FMA (12 flop per row-code pair); the HBM
        # bytes are negligible (rows x 40 B + a 256 KB codebook), so the honest bound is the fp32 ALU.
        rows = 32 * 32 * 32
        k6_ms = hv.bench_vq_argmin(rows, 20)
        k6_bytes = rows * (16 + 8 + 16) + 16384 * 16
        ops = rows * 16384 * 12.0
        alu_peak = 148 * 128 * 1.965e9   # fp32 lanes x clock: separately rounded mul / add (no FMA contraction)
        line["roofline_k6"] = {"bound": "hbm", "kernel": "vq_argmin_kernel (+ code norms, z/scale)", "rows": rows, "codes": 16384,
                               "achieved": k6_bytes / (k6_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                               "frac": k6_bytes / (k6_ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                               "bytes_per_launch": k6_bytes, "ms_per_launch": k6_ms,
                               "alu": {"flop_per_launch": ops, "achieved_tflops": ops / (k6_ms * 1e-3) / 1e12,
                                       "peak_tflops_fp32_nofma": alu_peak / 1e12,
                                       "frac": ops / (k6_ms * 1e-3) / alu_peak,
                                       "note": "the kernel is ALU-bound: HBM frac is small by construction"}}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of gn_stats16x8_kernel / gn_apply16x8_kernel at the shape above,
# from one `ncu --set full` capture (profiles/r2_ncu_gn_kernels.csv): 134.24 + 5.71 MB and 134.26 + 92.1 MB (the
# rest of the 134 MB the apply pass writes is still in L2 when the kernel ends)
K2_NCU_TRAFFIC = {"stats": 139.95e6, "apply": 226.4e6, "source": "profiles/r2_ncu_gn_kernels.csv"}
# DRAM bytes of all implicit-GEMM launches of ONE CFG UNet step at 64 images (177 + 2 launches), ncu cold-cache replays:
# profiles/r2_launches_unet_step_b64_summary.txt (16009 + 3973 + 957 + 2203 + 99 MB)
GEMM_DRAM_BYTES_PER_UNET_STEP = {(32, 64): 23.24e9}


if __name__ == "__main__":
    main()
