"""Multi-GPU plumbing for the trace path: the tree is built once and replicated, rays are sharded by
interleaved image rows, the framebuffer is reassembled with an all-gather.  torch.distributed does the transport (NCCL on GPUs;
the same code runs over gloo with CPU tensors in the tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .tree import Tree


def my_rows(height_per_rank: int, rank: int, world: int) -> np.ndarray:
    """global row indices rendered by `rank`: rank, rank + world, ... (height_per_rank of them)"""
    return rank + world * np.arange(height_per_rank)


def gather_framebuffer(pixels_local: torch.Tensor, width: int, gathered: torch.Tensor | None = None) -> torch.Tensor:
    """all-gather the per-rank framebuffers ((H*W,) int32 each, rows rank, rank+N, ...) and return the (N*H, W) image in
    global row order on every rank.  `gathered` ((N, H*W)) may be preallocated."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return pixels_local.view(-1, width)
    n = pixels_local.numel()
    if gathered is None:
        gathered = torch.empty((world, n), dtype=pixels_local.dtype, device=pixels_local.device)
    dist.all_gather_into_tensor(gathered.view(-1), pixels_local.contiguous())
    h = n // width
    # gathered[r, k*W:(k+1)*W] is global row k*world + r  ->  (h, world, W) row-major == global order
    return gathered.view(world, h, width).permute(1, 0, 2).reshape(h * world, width)


def replicate_tree(tree: Tree | None, device, src: int = 0) -> Tree:
    """Build-once replication: rank `src` passes its built Tree, every other rank passes None; returns a traceable Tree on every
    rank (the source's own tree on `src`).  Transport: one broadcast of the sizes, then one broadcast per array."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return tree
    L = _lib.load()
    rank = dist.get_rank()
    sizes = torch.zeros(12, dtype=torch.int64, device=device)
    world_box = torch.zeros(7, dtype=torch.float32, device=device)
    if rank == src:
        s = (C.c_uint64 * 8)()
        _lib.check(L.bmTreeArraySizes(tree.h, s), "bmTreeArraySizes")
        wmin, wmax, min_leaf, max_depth = tree.world()
        layout = tree.node_layout()
        sizes.copy_(torch.tensor([int(x) for x in s] + [tree.root(), 1 if layout != _lib.BM_NODES_BINARY else 0, max_depth, 0], dtype=torch.int64))
        world_box.copy_(torch.tensor(list(wmin) + list(wmax) + [min_leaf], dtype=torch.float32))
    dist.broadcast(sizes, src)
    dist.broadcast(world_box, src)
    sz = [int(x) for x in sizes.cpu()]
    wb = world_box.cpu().numpy().astype(np.float32)
    bufs = [torch.empty(max(b, 16), dtype=torch.uint8, device=device) for b in sz[:5]]
    if rank == src:
        _lib.check(L.bmTreeExport(tree.h, *[b.data_ptr() for b in bufs]), "bmTreeExport")
        tree.stream.synchronize()
    for b in bufs:
        dist.broadcast(b, src)
    if rank == src:
        return tree
    out = Tree.__new__(Tree)
    out.L = L; out.device = torch.device(device); out._meshes = []
    h = C.c_void_p()
    wmin = np.ascontiguousarray(wb[0:3]); wmax = np.ascontiguousarray(wb[3:6])
    _lib.check(L.bmTreeImport(C.byref(h), *[b.data_ptr() for b in bufs], sz[5], sz[6], sz[7], sz[8], sz[9], wmin.ctypes.data, wmax.ctypes.data, float(wb[6]), sz[10]), "bmTreeImport")
    out.h = h
    out._imported_buffers = bufs          # keep the adopted memory alive with the handle
    out.use_stream(torch.cuda.current_stream(out.device))
    return out
