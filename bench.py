#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: 256x256 images/s, 50-step DDIM + CFG).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

One "step" = one full sampling job of the per-GPU batch: 50 DDIM steps (each = CFG-doubled UNet
+ fused CFG/DDIM update) + KL decode to 256x256 (+ NCCL all-gather of the images when N > 1).
Weak scaling: 8 images per GPU, so N = 8 is BASELINE.json configs[2] (latent [64,32,32,4]).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic work (SURVEY 8d / BASELINE.md section 2), 2*MAC, unpadded dims
GFLOP_UNET_STEP_PER_IMAGE = 364.15      # one CFG step (2 UNet passes), 32x32 latent
GFLOP_CTX_KV_PER_IMAGE = 9.84           # loop-invariant context K/V projections, hoisted
GEMM_DRAM_BYTES_PER_UNET_STEP_B8 = 3.904e9  # profiles/r1_dram_unet_step.csv (3893.3 MB read + 10.4 MB written, 179 launches)
GFLOP_KL_DECODE_PER_IMAGE = 622.19
METRIC = "images_per_s_256x256_ddim50_cfg"
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


# --------------------------------------------------------------------------------------
# CPU arm: the NumPy restatement of the reference (oracle/), the only place bench.py runs it
# --------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, quiet=False):
    from oracle import ldm_oracle as O
    cfg = O.FULL_CONFIG
    us = O.unet_spec(cfg["unet"])
    Wu = O.as_dict(us, O.init_weights(us, 0))
    as_ = O.ae_spec(cfg["autoencoder_kl"], "kl")
    Wa = O.as_dict(as_, O.init_weights(as_, 2))
    sched = O.ddim_schedule(**cfg["ldm"])
    rng = np.random.default_rng(1234)
    xt = rng.standard_normal((1, 32, 32, 4), dtype=np.float32)
    ctx = np.random.default_rng(3).standard_normal((2, 77, 1280), dtype=np.float32)

    def one_step(x, index):
        t = np.full([2], sched["ddim_steps"][index], np.int32)
        e = O.unet_forward(Wu, cfg["unet"], np.concatenate([x, x]), t, ctx)
        return O.ddim_update(x, e[:1], e[1:], None, O.ddim_coeffs(sched, index), 5.0)[0]

    t0 = time.perf_counter()
    O.decode_first_stage(Wa, cfg["autoencoder_kl"], "kl", xt * np.float32(0.18215))
    t_dec = time.perf_counter() - t0
    for i in range(warmup):
        xt = one_step(xt, 49 - i % 50)
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        xt = one_step(xt, 49 - (warmup + i) % 50)
        ts.append(time.perf_counter() - t0)
    t_step = float(np.mean(ts))
    per_image = 50 * t_step + t_dec
    return dict(value=1.0 / per_image, t_step=t_step, t_dec=t_dec, cores=os.cpu_count(),
                sample=f"B=1: {steps} CFG UNet steps ({t_step:.2f} s each) + 1 KL decode ({t_dec:.2f} s), "
                       f"extrapolated to 50 steps + decode per image; NumPy fp32 + OpenBLAS, all host threads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=8)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch_per_gpu
    workload = (f"txt2img-f8-large random-init, latent [{B * world},32,32,4] ({B}/GPU), {args.ddim_steps} DDIM steps "
                f"eta=0 + CFG (guidance 5), KL decode to 256x256" + (", NCCL all-gather of images" if world > 1 else ""))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(max(args.steps, 1), min(args.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "CPU restatement of the reference (oracle/, NumPy fp32); "
                       "TensorFlow is not installable here, so this is kind=port, not the TF2 sampler itself"},
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
    from ldm_tf2_b200.sampler import (AutoencoderKL, LatentDiffusionModelSampler, TransformerModel, UNet)

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = synth.FULL_CONFIG
    precision = os.environ.get("LDM_B200_PRECISION", lib.DEFAULT_PRECISION)
    ldm_kw = dict(cfg["ldm"])
    ldm_kw["num_ddim_steps"] = args.ddim_steps
    # public API objects, exactly as run_ldm_sampler.py:56-83 builds them
    text = TransformerModel(**cfg["cond_stage_model"])
    unet = UNet(**cfg["unet"])
    ae = AutoencoderKL(**{k: v for k, v in cfg["autoencoder_kl"].items()})
    sampler = LatentDiffusionModelSampler(unet, ae, text, device=local_rank, **ldm_kw)
    h = sampler.handle
    for model, seed in ((h.TEXT, 1), (h.UNET, 0), (h.AE, 2)):
        h.set_weights(model, synth.random_weights(h, model, seed))
    h.finalize()
    h.configure_sampler(sampler.schedule.ddim_steps, sampler.schedule.coeff_table())

    ids = tokens.default_token_ids(B)
    # global seeded x_T, sliced per rank: results do not depend on the GPU count
    xg = np.random.default_rng(1234).standard_normal((B * world, 32, 32, 4), dtype=np.float32)
    x_host = torch.empty((B, 32, 32, 4), dtype=torch.float32, pin_memory=True)
    x_host.copy_(torch.from_numpy(parallel.shard_batch(xg, rank, world)))
    x_np = x_host.numpy()
    dev = torch.device("cuda", local_rank)
    x_dev = x_host.to(dev)
    lat_dev = torch.empty_like(x_dev)
    img_dev = torch.empty((B, 256, 256, 3), dtype=torch.float32, device=dev)
    ctx = h.encode_text(ids)
    h.set_context(ctx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def device_step():
        """inputs resident in HBM: sample + decode (+ all-gather); returns device ms"""
        lib.check(h.lib.ldm_sample(h._h, lib.ptr(x_dev.data_ptr()), None, B, 32, 32, 5.0, lib.ptr(lat_dev.data_ptr()),
                                   None, 0, 1))
        lib.check(h.lib.ldm_decode(h._h, lib.ptr(lat_dev.data_ptr()), B, 32, 32, 0.18215,
                                   lib.ptr(img_dev.data_ptr()), None))
        t = h.timing()
        ms = t["loop_ms"] + t["decode_ms"]
        if world > 1:
            ev0.record()
            parallel.allgather_images(img_dev, B * world)
            ev1.record()
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
        return ms, t

    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = h.timing()["launches"]
    w0 = time.perf_counter()
    dev_ms, loop_ms, dec_ms = 0.0, 0.0, 0.0
    for _ in range(args.steps):
        ms, t = device_step()
        dev_ms += ms
        loop_ms += t["loop_ms"]
        dec_ms += t["decode_ms"]
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    launches = h.timing()["launches"] - l0
    clk = clocks.stop()

    # end to end through the public API: host x_T in, host images out, every call
    def e2e_step():
        return sampler.ddim_p_sample_loop(ids, (B, 32, 32, 4), 5.0, x_init=x_np)

    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            images = e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - e0) * 1e3
    h2d = x_np.nbytes + ids.nbytes + ctx.nbytes
    d2h = images.nbytes + ctx.nbytes + x_np.nbytes  # images + context + final latents read back

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_ms, wall_ms, e2e_ms = allmax(dev_ms), allmax(wall_ms), allmax(e2e_ms)
    launches_total = launches
    if world > 1:
        t = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        launches_total = int(t.item())

    line = None
    if rank == 0:
        peaks = measured_peaks()
        prof = h.profile_unet_step(B, 32, 32, 3)
        step_gflop = B * (GFLOP_UNET_STEP_PER_IMAGE - GFLOP_CTX_KV_PER_IMAGE)
        # Time of the GEMM kernel inside the replayed step graph = full step - the same graph without
        # its GEMM launches (CUDA events around 20 graph replays each).  Events around every single
        # launch of an eager step (prof) add ~4 us of launch latency to each of the 179 launches and
        # overstate the kernel's share (81 % against ncu's 62 %); that figure is kept as a cross-check.
        t_full = h.bench_unet_step(B, 32, 32, 20, True)
        t_rest = h.bench_unet_step(B, 32, 32, 20, True, skip_gemm=True)
        gemm_ms = t_full - t_rest
        achieved = step_gflop / gemm_ms  # GFLOP/ms == TFLOP/s
        k5_ms = h.bench_ddim_update(B, 32, 32, False, 200)
        k5_bytes = 4 * 4 * B * 32 * 32 * 4  # 3 reads + 1 write of fp32 [B,32,32,4]
        # K2 GroupNorm at the decoder's largest activation [B, 256*256, 128] (HBM-resident: 268 MB fp32)
        gn_n, gn_hw, gn_c = B, 256 * 256, 128
        gn_stats_ms, gn_apply_ms = h.bench_groupnorm(gn_n, gn_hw, gn_c, 10)
        gn_el = gn_n * gn_hw * gn_c
        total_images = B * world * args.steps
        value = total_images / (dev_ms / 1e3)
        gflop_per_image = args.ddim_steps * GFLOP_UNET_STEP_PER_IMAGE + GFLOP_KL_DECODE_PER_IMAGE
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic",
            "config": {"workload": workload, "operands": f"{precision} tensor-core operands, fp32 accumulate, fp32 residual stream", "global_batch": B * world, "parallelism": f"dp{world} (sample-sharded replicas)",
                       "l2": "not flushed explicitly: 1.75 GB of bf16 weights stream through L2 every UNet step",
                       "timing": "CUDA events on the library stream (+ torch events for the all-gather), max over ranks"},
            "ms_per_unet_step": loop_ms / args.steps / args.ddim_steps,
            "ms_decode": dec_ms / args.steps,
            "wall_ms_per_step": wall_ms / args.steps,
            "model_tflops": value * gflop_per_image / 1e3 / world,
            "e2e": {"value": total_images / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "api": "LatentDiffusionModelSampler.ddim_p_sample_loop(ids, shape, guidance) incl. text encoder"},
            "gpu_launches": int(launches_total),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "implicit_gemm_kernel (tcgen05)", "achieved": achieved,
                         "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                         # DRAM bytes of the 179 GEMM launches of one UNet step (ncu, cold-cache replays):
                         # profiles/r1_dram_unet_step.csv; only valid for the default 8-image workload
                         "traffic": GEMM_DRAM_BYTES_PER_UNET_STEP_B8 if B == 8 else None,
                         "traffic_unit": "bytes per UNet step (all GEMM launches)",
                         "peak_source": peaks["source"] + ", sustained bf16",
                         "launches_per_unet_step": prof["gemm_launches_per_step"],
                         "kernel_ms_per_unet_step": gemm_ms, "step_ms_graph": t_full, "step_ms_graph_without_gemm": t_rest,
                         "kernel_share_of_step": gemm_ms / t_full,
                         "kernel_ms_event_per_launch_sum": prof["gemm_ms_per_step"], "eager_step_ms": prof["step_ms"],
                         "achieved_event_per_launch": step_gflop / prof["gemm_ms_per_step"],
                         "algorithmic_gflop_per_unet_step": step_gflop},
            "roofline_k5": {"bound": "hbm", "kernel": "ddim_update_kernel", "achieved": k5_bytes / (k5_ms * 1e-3) / 1e9,
                            "peak": peaks["hbm"], "unit": "GB/s", "frac": k5_bytes / (k5_ms * 1e-3) / 1e9 / peaks["hbm"],
                            "traffic": None, "bytes_per_launch": k5_bytes, "ms_per_launch": k5_ms},
        }
        line["roofline_k2"] = {
            "bound": "hbm", "kernel": "gn_stats_kernel + gn_apply_kernel (GroupNorm(32)+SiLU -> 16-bit operand)",
            "shape": [gn_n, gn_hw, gn_c], "unit": "GB/s", "peak": peaks["hbm"],
            "stats": {"bytes_per_launch": gn_el * 4, "ms_per_launch": gn_stats_ms,
                      "achieved": gn_el * 4 / (gn_stats_ms * 1e-3) / 1e9,
                      "frac": gn_el * 4 / (gn_stats_ms * 1e-3) / 1e9 / peaks["hbm"]},
            "apply": {"bytes_per_launch": gn_el * 6, "ms_per_launch": gn_apply_ms,
                      "achieved": gn_el * 6 / (gn_apply_ms * 1e-3) / 1e9,
                      "frac": gn_el * 6 / (gn_apply_ms * 1e-3) / 1e9 / peaks["hbm"]},
            "traffic": None}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(2, 1)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        else:
            line["cpu_baseline"] = None
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sampler.close()


if __name__ == "__main__":
    main()
