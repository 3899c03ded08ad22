"""CPU tier, world_size 2 over gloo: the multi-GPU sharding logic (interleaved rows per rank -> all-gather -> reassembled frame).
The tracer in this test is the ORACLE (there is no GPU here); what is under test is the host-side sharding / reassembly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, W, H, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.bindings import OracleLib, OracleScene
    from raytracercuda_b200 import synth, workloads
    from raytracercuda_b200.dist import gather_framebuffer, my_rows
    wl = workloads.get("c2small")
    O = OracleLib()
    s = OracleScene(O); s.add_mesh(wl.positions, wl.indices, wl.normals); s.build()
    fr = workloads.interleaved_rows_frustum(wl.frustum, H, rank, world)
    rays = O.initial_rays(W, H, fr)
    # the rank's table must be exactly rows rank, rank+world, ... of the full-height table
    full = O.initial_rays(W, H * world, wl.frustum).reshape(H * world, W, 3)
    mine = full[my_rows(H, rank, world)].reshape(-1, 3)
    assert np.abs(mine - rays).max() < 2e-6, np.abs(mine - rays).max()
    r = s.trace_camera(rays, wl.eye, synth.IDENTITY, 0)
    local = torch.from_numpy(r["pixels"].view(np.int32).copy())
    frame = gather_framebuffer(local, W)
    assert frame.shape == (H * world, W)
    # reference: the whole frame traced by one process with the per-rank tables stitched (rounding-identical inputs)
    if rank == 0:
        want = np.zeros((H * world, W), np.uint32)
        for q in range(world):
            rq = O.initial_rays(W, H, workloads.interleaved_rows_frustum(wl.frustum, H, q, world))
            want[my_rows(H, q, world)] = s.trace_camera(rq, wl.eye, synth.IDENTITY, 0)["pixels"].reshape(H, W)
        got = frame.numpy().view(np.uint32)
        assert np.array_equal(got, want)
        assert (got != 255 << 8).sum() > 500
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_interleaved_rows_allgather_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 96, 40, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


def test_row_partition_is_exact():
    sys.path.insert(0, ROOT)
    from raytracercuda_b200.dist import my_rows
    for world in (1, 2, 4, 8):
        rows = np.sort(np.concatenate([my_rows(135, r, world) for r in range(world)]))
        assert np.array_equal(rows, np.arange(135 * world))
