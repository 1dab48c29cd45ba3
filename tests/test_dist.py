"""N > 1 host logic on CPU: two gloo ranks render interleaved tiles of one frame and reduce it.
The renderer behind render_frame() here is the CPU harness (the oracle); on the GPU box the same
function drives rtw_render_device + NCCL (bench.py, tests/test_gpu_render.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import raytracer_weekend_b200 as rtw
from raytracer_weekend_b200 import dist as rdist
from conftest import Oracle

W, H, SPP = 96, 64, 3


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        with rtw.Scene.from_name(orc, "cornell-box", W / H, seed=1) as s:
            cam = s.cameras[0]
            params = s.params(W, H, SPP, seed=4)
            accum = torch.zeros(H * W * 3, dtype=torch.float32)

            def render_into(p, buf):
                a, st = s.render(cam, p)
                buf.copy_(torch.from_numpy(a.reshape(-1)))
                return st

            st = rdist.render_frame(render_into, params, accum, tile_size=16)
            seg = torch.tensor([st.segments], dtype=torch.int64)
            dist.all_reduce(seg)
            if rank == 0:
                np.savez(out_path, accum=accum.numpy(), segments=seg.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_partition_and_reduce(tmp_path, oracle, world):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    with rtw.Scene.from_name(oracle, "cornell-box", W / H, seed=1) as s:
        full, st = s.render(s.cameras[0], s.params(W, H, SPP, seed=4))
    # each pixel is summed by exactly one rank and the stream is keyed by (pixel, sample):
    # the merged frame is bit-identical to the single-process frame
    assert np.array_equal(got["accum"].view(np.uint32), full.reshape(-1).view(np.uint32))
    assert int(got["segments"][0]) == st.segments


def test_partition_helper():
    p = rtw.RenderParams(width=100, height=50, spp=4)
    q = rdist.partition(p, 2, 5, tile_size=16)
    assert (q.part_rank, q.part_count, q.tile_size) == (2, 5, 16) and p.part_count == 0
