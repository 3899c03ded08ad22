"""does a tile-major (8x4 pixel tiles) order of the bounce-ray list trace faster than the row-major list?"""
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import _lib, host, workloads
from raytracercuda_b200.tree import Tree, make_mesh_table
L = _lib.load(); dev = torch.device("cuda:0")
wl = workloads.get("c2"); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); cam.setInitialRays(W, H, *wl.frustum)
rays = torch.from_numpy(cam.initialRays()).to(dev)
Pd = torch.from_numpy(wl.positions).to(dev); Id = torch.from_numpy(wl.indices.view(np.int32)).to(dev).reshape(-1)
tree = Tree(dev); tree.add_mesh(Pd, Id); tree.build()
hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
rays2 = torch.empty((n, 6), dtype=torch.float32, device=dev); src2 = torch.empty(n, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
hits2 = torch.empty((n, 4), dtype=torch.float32, device=dev)
f3 = (C.c_float * 3)(0.0, 0.0, -2.1); f9 = (C.c_float * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)
_lib.check(L.bmTraceCamera(tree.h, rays.data_ptr(), n, W, f3, f9, 0, hits.data_ptr(), None, None, None), "a")
_lib.check(L.bmBounceRaysCamera(tree.h, rays.data_ptr(), f3, f9, hits.data_ptr(), n, 0, 1, 0, rays2.data_ptr(), src2.data_ptr(), cnt.data_ptr()), "b")
torch.cuda.synchronize(); nb = int(cnt.cpu()[0])
src = src2[:nb].long(); px = src % W; py = src // W
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
def time_list(R, label):
    ms = []
    for _ in range(8):
        flush.zero_(); e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record()
        _lib.check(L.bmTraceRays(tree.h, R.data_ptr(), nb, 0, hits2.data_ptr(), None), "t"); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print("%-28s %.4f ms (min %.4f)" % (label, sum(ms[2:]) / len(ms[2:]), min(ms)), flush=True)
base = rays2[:nb].contiguous()
time_list(base, "row-major (current)")
for tw, th in ((8, 4), (16, 8), (32, 16), (4, 4)):
    tiles_x = (W + tw - 1) // tw
    key = ((py // th) * tiles_x + (px // tw)) * (tw * th) + (py % th) * tw + (px % tw)
    perm = torch.argsort(key, stable=True)
    time_list(base[perm].contiguous(), "tile-major %dx%d" % (tw, th))
# Morton order of the pixels
def part1by1(v):
    v = v & 0xFFFF; v = (v | (v << 8)) & 0x00FF00FF; v = (v | (v << 4)) & 0x0F0F0F0F; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v
key = part1by1(px) | (part1by1(py) << 1)
time_list(base[torch.argsort(key, stable=True)].contiguous(), "Morton")
