"""Multi-GPU host path: the batch of samples shards across ranks (one process per GPU,
full weight replica each, both CFG halves of an image on the same rank), there is no collective
inside the DDIM loop, and ONE all-gather collects the decoded images.

On GPUs the all-gather is `ldm_allgather_images` of the library (NCCL over NVLink on the handle's stream,
comm.cu); the 128-byte NCCL id travels from rank 0 to the other ranks over a tiny TCP rendezvous on
MASTER_ADDR / MASTER_PORT+1.. (`exchange_bytes`) -- no PyTorch anywhere on this path.  The torch.distributed
variant (`allgather_images`) remains for the CPU tests, which run the same shard / pad logic on gloo."""
from __future__ import annotations

import os
import socket
import struct
import time

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous block partition of `total` images: rank r gets [lo, hi).  The remainder goes to
    the first ranks, so any total works (ragged shards are padded in allgather_images)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(global_array, rank: int, world: int, axis: int = 0):
    """Slice of a globally seeded tensor (x_T [B,h,w,4], or noise [S,B,h,w,4] with axis=1): every
    rank draws from the same global tensor, so results do not depend on the GPU count."""
    lo, hi = shard_range(global_array.shape[axis], rank, world)
    idx = [slice(None)] * global_array.ndim
    idx[axis] = slice(lo, hi)
    return np.ascontiguousarray(global_array[tuple(idx)])


def shard_token_ids(token_ids, rank: int, world: int):
    """get_token_ids layout is B uncond rows then B cond rows (run_ldm_sampler.py:42-45); a rank
    needs the uncond AND cond rows of its own images."""
    b = token_ids.shape[0] // 2
    lo, hi = shard_range(b, rank, world)
    return np.concatenate([token_ids[lo:hi], token_ids[b + lo:b + hi]], axis=0)


def allgather_images(local, total: int, group=None):
    """All ranks end up with the [total, H, W, 3] image tensor in global sample order.  `local` is
    a torch tensor (CUDA for NCCL, CPU for gloo) holding this rank's shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < maxn:  # ragged last shards: pad to a common size for the collective
        pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    out = torch.empty((world * maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(hi - lo == maxn for lo, hi in sizes):
        return out
    parts = [out[r * maxn: r * maxn + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


# ------------------------------------------------------------------------------------------------
# NCCL inside the library: rendezvous + gather (no torch)
# ------------------------------------------------------------------------------------------------
_MAGIC = b"LDMB200\x01"
_PORT_TRIES = 8   # rank 0 binds the first free port of MASTER_PORT+1 .. +8; the other ranks probe the same list


def _job_token(addr: str, base_port: int, world: int) -> bytes:
    import hashlib
    return hashlib.sha256(f"{addr}:{base_port}:{world}".encode()).digest()[:8]


def _recv_exact(c, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        chunk = c.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("rendezvous closed early")
        buf += chunk
    return buf


def exchange_bytes(payload, rank: int, world: int, addr: str = None, port: int = None, timeout: float = 120.0) -> bytes:
    """Rank 0 serves `payload` to the world - 1 other ranks over TCP; every rank returns it.

    The first port is MASTER_PORT + 1 (or `port`).  When something else already listens there, rank 0 moves on to
    the next of `_PORT_TRIES` consecutive ports and the other ranks probe the same list: a connection counts only
    after a hello that carries the magic and this job's token (address, base port, world size), so a foreign service
    or another job's rendezvous on a neighbouring port is skipped instead of believed."""
    if world == 1:
        return payload
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    base = int(port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + 1)
    ports = [base + i for i in range(_PORT_TRIES)]
    hello = _MAGIC + _job_token(addr, base, world)
    deadline = time.time() + timeout
    if rank == 0:
        srv, err = None, None
        for p in ports:
            s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            try:
                s.bind((addr, p))
                s.listen(world)
                srv = s
                break
            except OSError as e:
                err = e
                s.close()
        if srv is None:
            raise OSError(f"rendezvous: no free port in {ports[0]}..{ports[-1]} on {addr}: {err}")
        try:
            served = 0
            while served < world - 1:
                left = deadline - time.time()
                if left <= 0:
                    raise TimeoutError(f"rendezvous: {served} of {world - 1} ranks arrived within {timeout:.0f} s")
                srv.settimeout(left)
                try:
                    c, _ = srv.accept()
                except socket.timeout:
                    continue
                with c:
                    c.settimeout(5.0)
                    try:
                        if _recv_exact(c, len(hello)) != hello:
                            continue   # not one of ours
                        c.sendall(_MAGIC + struct.pack("<I", len(payload)) + payload)
                        served += 1
                    except (ConnectionError, socket.timeout, OSError):
                        continue
        finally:
            srv.close()
        return payload
    i = 0
    while True:
        p = ports[i % len(ports)]
        i += 1
        try:
            with socket.create_connection((addr, p), timeout=5.0) as c:
                c.settimeout(5.0)
                c.sendall(hello)
                if _recv_exact(c, len(_MAGIC)) != _MAGIC:
                    raise ConnectionError("not the rendezvous")
                (n,) = struct.unpack("<I", _recv_exact(c, 4))
                return _recv_exact(c, n)
        except (ConnectionRefusedError, ConnectionError, socket.timeout, OSError):
            if time.time() > deadline:
                raise
            if i % len(ports) == 0:
                time.sleep(0.05)


def init_comm(handle, rank: int, world: int, addr: str = None, port: int = None):
    """ncclGetUniqueId on rank 0 -> TCP rendezvous -> ncclCommInitRank on every rank's handle."""
    from . import lib
    nccl = lib.nccl_library_path()
    uid = lib.comm_unique_id(nccl) if rank == 0 else None
    uid = exchange_bytes(uid, rank, world, addr, port)
    handle.comm_init(uid, rank, world, nccl)
    return handle


def gather_plan(total: int, world: int, per_image: int):
    """(largest shard in images, per-rank [lo, hi) list): ranks pad their shard to the largest one."""
    sizes = [shard_range(total, r, world) for r in range(world)]
    return max(hi - lo for lo, hi in sizes), sizes


def allgather_images_nccl(handle, local, total: int, out=None):
    """local: numpy [b_local, H, W, 3] float32 (host) -> numpy [total, H, W, 3] on every rank, global sample
    order; or lib.DevPtr in / DevPtr out (device-resident, equal shards only)."""
    from . import lib
    world = handle_world(handle)
    if isinstance(local, lib.DevPtr):
        b = local.shape[0]
        if b * world != total:
            raise ValueError("device-resident gather needs equal shards")
        handle.allgather(local, local.size, out)
        return out
    local = np.ascontiguousarray(local, dtype=np.float32)
    per = int(np.prod(local.shape[1:]))
    maxn, sizes = gather_plan(total, world, per)
    send = local
    if local.shape[0] < maxn:
        send = np.zeros((maxn,) + local.shape[1:], np.float32)
        send[: local.shape[0]] = local
    buf = np.empty((world * maxn,) + local.shape[1:], np.float32)
    handle.allgather(send, maxn * per, buf)
    if all(hi - lo == maxn for lo, hi in sizes):
        return buf
    return np.concatenate([buf[r * maxn: r * maxn + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], axis=0)


def handle_world(handle) -> int:
    return getattr(handle, "_world", 1)
