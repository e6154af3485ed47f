"""SURVEY 8(e) on real GPUs: samples sharded over N ranks (one process per GPU, full weight replica each), no
collective inside the loop, ONE all-gather of the decoded images through the library's own NCCL communicator
(ldm_allgather_images, comm.cu; the NCCL id travels over the TCP rendezvous of ldm_tf2_b200/parallel.py).

The check is the reference's "multi-GPU invariance" (SURVEY 4): the gathered tensor must equal, bit for bit, what
one GPU computes for every shard of the same globally seeded x_T (bench.py --verify recomputes each shard on rank 0
with the same per-GPU batch, i.e. the same kernels).  Needs >= 2 GPUs: skipped on the single-GPU test box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q` (output kept in profiles/)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_two_rank_gather_is_bitwise_the_single_gpu_result():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--weak", "4", "--verify", "--no-rooflines", "--no-cpu-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=1200)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    print(json.dumps({k: line[k] for k in ("value", "n_gpus", "images_sha256_16", "verify", "scaling")}))
    assert line["n_gpus"] == 2 and line["config"]["global_batch"] == 8
    assert line["verify"] == {"gathered_equals_local_recompute_bitwise": True, "shards": 2}
