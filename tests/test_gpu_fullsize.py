"""GPU tier: BASELINE config 2 at its FULL size (327,680-triangle stand-in, 1920x1080 primary rays + one diffuse bounce per hit) through
size-independent properties, plus a strided sample of the same frame against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.bindings import OracleScene

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def _geometry_consistent(torch, o, d, hits, soup_v0, soup_e1, soup_e2, what):
    """every reported hit lies on its triangle: o + t d == v0 + u e1 + v e2 (to rounding), 0 <= u, v, u + v <= 1"""
    tri = hits[:, 3].view(torch.int32).long()
    hit = tri >= 0
    assert int(hit.sum()) > 0, what
    t, u, v = hits[hit, 0:1].double(), hits[hit, 1:2].double(), hits[hit, 2:3].double()
    k = tri[hit]
    p_ray = o[hit].double() + t * d[hit].double()
    p_tri = soup_v0[k].double() + u * soup_e1[k].double() + v * soup_e2[k].double()
    err = (p_ray - p_tri).abs().max(dim=1).values
    scale = 1.0 + t.abs().squeeze(1)
    rel = err / scale
    # float32 Moeller-Trumbore without an epsilon (the reference's arithmetic) is ill-conditioned on grazing triangles: the bulk sits at
    # rounding level, the worst few at 1e-4
    assert float(rel.max()) < 1e-3 and float(rel.median()) < 2e-6 and float(torch.quantile(rel[:: max(1, rel.numel() // 200000)], 0.999)) < 5e-5, \
        "%s: hit points off their triangles: max %g median %g" % (what, float(rel.max()), float(rel.median()))
    assert bool((u >= 0).all()) and bool((v >= 0).all()) and bool((u + v <= 1.0).all()), what
    return hit


def test_config2_full_frame_properties(oracle_lib, beam_lib, cuda_device):
    import torch
    from raytracercuda_b200 import _lib, workloads
    from raytracercuda_b200.host import ICamera
    from raytracercuda_b200.tree import Tree, hits_to_numpy
    L = beam_lib
    dev = cuda_device
    wl = workloads.get("c2")
    W, H = wl.width, wl.height
    n = W * H
    assert wl.num_tris == 327680 and n == 2073600
    Pd = torch.from_numpy(wl.positions).to(dev); Id = torch.from_numpy(wl.indices.view(np.int32)).to(dev)
    t = Tree(dev); t.add_mesh(Pd, Id.reshape(-1)); t.build()
    st = t.stats()
    assert st["triangles"] == wl.num_tris and st["nodes"] == st["leaves"] - 1 and st["depth"] == 31 and st["overflow_leaves"] == 0
    cam = ICamera.create(); assert cam.setInitialRays(W, H, *wl.frustum) == 0
    table = torch.from_numpy(cam.initialRays()).to(dev)
    eye = (C.c_float * 3)(*wl.eye); ori = (C.c_float * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)
    hits = [torch.empty((n, 4), dtype=torch.float32, device=dev) for _ in range(3)]
    for k in range(3):                                   # frame 0 in row order, frames 1-2 under the tile schedule
        _lib.check(L.bmTraceCamera(t.h, table.data_ptr(), n, W, eye, ori, 0, hits[k].data_ptr(), None, None, None), "bmTraceCamera")
    torch.cuda.synchronize()
    assert torch.equal(hits[0].view(torch.int32), hits[1].view(torch.int32)) and torch.equal(hits[0].view(torch.int32), hits[2].view(torch.int32))
    V = Pd[Id.long()]                                    # (F, 3, 3)
    v0, e1, e2 = V[:, 0], V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]
    o = torch.tensor(wl.eye, dtype=torch.float32, device=dev).expand(n, 3)
    hit = _geometry_consistent(torch, o, table, hits[0], v0, e1, e2, "primary")
    nh = int(hit.sum())
    assert 400_000 < nh < 700_000
    miss = ~hit
    assert bool((hits[0][miss, 0] == torch.finfo(torch.float32).max).all())
    # one diffuse bounce per hit: the count is the number of hits, sources are exactly the hit pixels in order, and the bounce hits are
    # geometrically consistent too
    rays2 = torch.empty((n, 6), dtype=torch.float32, device=dev); src = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev); hits2 = torch.empty((n, 4), dtype=torch.float32, device=dev)
    _lib.check(L.bmBounceRaysCamera(t.h, table.data_ptr(), eye, ori, hits[0].data_ptr(), n, 0, 1, 0, rays2.data_ptr(), src.data_ptr(), cnt.data_ptr()), "bmBounceRaysCamera")
    _lib.check(L.bmTraceRaysIndirect(t.h, rays2.data_ptr(), cnt.data_ptr(), n, 1, hits2.data_ptr(), None), "bmTraceRaysIndirect")      # bounce rays: closest hit, t >= 0
    torch.cuda.synchronize()
    assert int(cnt.cpu()[0]) == nh
    assert torch.equal(src[:nh].long(), torch.nonzero(hit).squeeze(1))
    dn = rays2[:nh, 3:].norm(dim=1)
    assert float((dn - 1).abs().max()) < 1e-4
    _geometry_consistent(torch, rays2[:nh, :3], rays2[:nh, 3:], hits2[:nh], v0, e1, e2, "bounce")
    # a strided sample of both batches against the oracle (identical ray buffers)
    s = OracleScene(oracle_lib); s.add_mesh(wl.positions, wl.indices, wl.normals); s.build()
    assert s.stats()["pairs"] == st["pairs"] and s.stats()["leaves"] == st["leaves"]
    sel = np.arange(0, n, 97)
    r = s.trace_camera(cam.initialRays()[sel], wl.eye, np.eye(3, dtype=np.float32), 0, NT)
    h = hits_to_numpy(hits[0][torch.from_numpy(sel).to(dev)])
    ok = r["tie"] == 0
    assert (h["tri"][ok] == r["tri"][ok]).mean() >= 0.9999
    both = ok & (r["tri"] != 0xFFFFFFFF) & (h["tri"] == r["tri"])
    assert np.array_equal(h["t"][both], r["t"][both])
    sel2 = np.arange(0, nh, 53)
    R = rays2[torch.from_numpy(sel2).to(dev)].cpu().numpy()
    r2 = s.trace_rays(R[:, :3], R[:, 3:], 1, NT)
    h2 = hits_to_numpy(hits2[torch.from_numpy(sel2).to(dev)])
    ok2 = r2["tie"] == 0
    assert (h2["tri"][ok2] == r2["tri"][ok2]).mean() >= 0.9999
    both2 = ok2 & (r2["tri"] != 0xFFFFFFFF) & (h2["tri"] == r2["tri"])
    rel = np.abs(h2["t"][both2] - r2["t"][both2]) / np.maximum(np.abs(r2["t"][both2]), 1e-20)
    assert rel.size and rel.max() <= 1e-5
    t.close(); s.close()
