import sys, os, numpy as np, torch, ctypes as C
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import _lib, host, workloads, synth
from raytracercuda_b200.tree import Tree, make_mesh_table
L = _lib.load(); dev = torch.device("cuda:0")
wl = workloads.get("c2"); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); cam.setInitialRays(W, H, *wl.frustum)
rays_pinned = torch.from_numpy(cam.initialRays()).pin_memory(); rays = rays_pinned.to(dev)
Pd = torch.from_numpy(wl.positions).to(dev); Id = torch.from_numpy(wl.indices.view(np.int32)).to(dev).reshape(-1); Nd = torch.from_numpy(wl.normals).to(dev)
table = make_mesh_table([(Pd, Id, Nd)], dev)
t = Tree(dev); t.add_mesh(Pd, Id); t.build()
hits = torch.empty((n, 4), device=dev); pix = torch.empty(n, dtype=torch.int32, device=dev)
f3 = (C.c_float * 3)(*wl.eye); f9 = (C.c_float * 9)(*np.eye(3, dtype=np.float32).reshape(-1))
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ms = []
    for _ in range(reps):
        e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), sum(ms) / len(ms)
for bands in (1, 2, 4, 8):
    nb = n // bands
    def prim():
        for k in range(bands):
            lo = k * nb
            L.bmTraceCamera(t.h, rays.data_ptr() + lo * 12, nb, W, f3, f9, 0, hits.data_ptr() + lo * 16, pix.data_ptr() + lo * 4, table.data_ptr(), None)
    print("primary trace in %d bands: min %.3f ms mean %.3f ms" % ((bands,) + timeit(prim)))
    for k in range(bands):
        lo = k * nb
        def one():
            L.bmTraceCamera(t.h, rays.data_ptr() + lo * 12, nb, W, f3, f9, 0, hits.data_ptr() + lo * 16, pix.data_ptr() + lo * 4, table.data_ptr(), None)
        print("   band %d: %.3f ms" % (k, timeit(one)[0]))
def h2d(): rays.copy_(rays_pinned, non_blocking=True)
print("H2D 25MB: %.3f ms" % timeit(h2d)[0])
ph = torch.empty(n, dtype=torch.int32).pin_memory()
def d2h(): ph.copy_(pix, non_blocking=True)
print("D2H 8MB: %.3f ms" % timeit(d2h)[0])
