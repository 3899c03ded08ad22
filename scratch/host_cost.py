import sys, os, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import _lib, host, workloads
from raytracercuda_b200.tree import Tree, make_mesh_table
L = _lib.load(); dev = torch.device("cuda:0")
wl = workloads.get("c2"); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); cam.setInitialRays(W, H, *wl.frustum)
rays = torch.from_numpy(cam.initialRays()).to(dev)
Pd = torch.from_numpy(wl.positions).to(dev); Id = torch.from_numpy(wl.indices.view(np.int32)).to(dev).reshape(-1); Nd = torch.from_numpy(wl.normals).to(dev)
tree = Tree(dev); tree.add_mesh(Pd, Id); tree.build()
table = make_mesh_table([(Pd, Id, Nd)], dev)
hits = torch.empty((n, 4), dtype=torch.float32, device=dev); pix = torch.empty(n, dtype=torch.int32, device=dev)
rays2 = torch.empty((n, 6), dtype=torch.float32, device=dev); src2 = torch.empty(n, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
hits2 = torch.empty((n, 4), dtype=torch.float32, device=dev)
f3 = lambda a: (C.c_float * 3)(*a); f9 = lambda a: (C.c_float * 9)(*a)
eye = (0.0, 0.0, -2.1); ori = (1, 0, 0, 0, 1, 0, 0, 0, 1)
def frame():
    _lib.check(L.bmTraceCamera(tree.h, rays.data_ptr(), n, W, f3(eye), f9(ori), 0, hits.data_ptr(), pix.data_ptr(), table.data_ptr(), None), "a")
    _lib.check(L.bmBounceRaysCamera(tree.h, rays.data_ptr(), f3(eye), f9(ori), hits.data_ptr(), n, 0, 1, 0, rays2.data_ptr(), src2.data_ptr(), cnt.data_ptr()), "b")
    _lib.check(L.bmTraceRaysIndirect(tree.h, rays2.data_ptr(), cnt.data_ptr(), n, 0, hits2.data_ptr(), None), "c")
for _ in range(5): frame()
torch.cuda.synchronize()
K = 200
t0 = time.perf_counter()
for _ in range(K): frame()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue per frame: %.1f us; wall per frame incl. GPU: %.1f us" % ((t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6))
