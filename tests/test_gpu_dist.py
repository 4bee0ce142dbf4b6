"""Multi-GPU correctness under pytest (tests/run_dist_gpu.py was never collected).

Spawns ``python -m torch.distributed.run`` with 2 ranks (one process per GPU, NCCL) on
tests/run_dist_gpu.py when >= 2 GPUs are visible: row-sharded 1:N search + all-gather + merge == numpy
oracle == 1-GPU search, and data-parallel detect == single-GPU detect.  Also a single-process,
two-device test of the C ABI (per-device launch state)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_search_and_dp_detect_on_two_gpus():
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "run_dist_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "run_dist_gpu ok: world=2" in r.stdout


def test_two_contexts_on_two_devices_in_one_process(capi, det_weights, rec_weights):
    """fr_create(device) must work for any device of the process: kernel attributes (opt-in dynamic
    shared memory) and the SM count are per device, not per process."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, (640, 640, 3), dtype=np.uint8) for _ in range(2)]
    crops = rng.integers(0, 256, (5, 112, 112, 3), dtype=np.uint8)
    q = rng.normal(size=(33, 512)).astype(np.float32)
    rows = rng.normal(size=(1000, 512)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    outs = []
    ctxs = [capi.Context(d, det_weights, rec_weights) for d in (1, 0)]     # device 1 FIRST
    for c in ctxs:
        dets = c.detect_batch(frames, 0.02, 0.4, cap=2048)                 # nms_kernel: 119 KB dynamic smem
        emb = c.embed_aligned(crops)
        g = capi.Gallery(c, 1000)
        g.add(rows)
        outs.append((dets, emb, g.search(q, 10)))
        g.close()
    (d0, e0, (s0, i0)), (d1, e1, (s1, i1)) = outs
    assert all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(d0, d1))
    assert np.array_equal(e0, e1) and np.array_equal(s0, s1) and np.array_equal(i0, i1)
    for c in ctxs:
        c.close()
