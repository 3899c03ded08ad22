"""scratch: the c2 frame's launches one by one (primary camera launch, bounce generation, bounce trace in closest mode) for ncu captures / A-B timing"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from raytracercuda_b200 import _lib, host, workloads
from raytracercuda_b200.tree import Tree, make_mesh_table
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c2"); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--bounce-mode", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0); L = _lib.load(); wl = workloads.get(a.workload); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); assert cam.setInitialRays(W, H, *wl.frustum) == 0
rays_dev = torch.from_numpy(cam.initialRays()).to(dev)
Pd = torch.from_numpy(wl.positions).to(dev); Id = torch.from_numpy(wl.indices.view(np.int32)).to(dev).reshape(-1); Nd = torch.from_numpy(wl.normals).to(dev)
table = make_mesh_table([(Pd, Id, Nd)], dev)
t = Tree(dev); t.add_mesh(Pd, Id); t.build(); print(t.stats(), flush=True)
f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v]); f9 = lambda m: (C.c_float * 9)(*[float(x) for x in np.asarray(m, np.float32).reshape(-1)])
eye, orient = wl.eye, np.eye(3, dtype=np.float32)
hits1 = torch.empty((n, 4), dtype=torch.float32, device=dev); pixels = torch.empty(n, dtype=torch.int32, device=dev)
rays2 = torch.empty((n, 6), dtype=torch.float32, device=dev); src2 = torch.empty(n, dtype=torch.int32, device=dev); cnt2 = torch.zeros(1, dtype=torch.int32, device=dev)
hits2 = torch.empty((n, 4), dtype=torch.float32, device=dev); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
nb64 = torch.zeros(1, dtype=torch.int64, device=dev)
for r in range(a.reps):
    flush.zero_(); e = [ev() for _ in range(4)]; torch.cuda.synchronize()
    e[0].record()
    _lib.check(L.bmTraceCamera(t.h, rays_dev.data_ptr(), n, W, f3(eye), f9(orient), 0, hits1.data_ptr(), pixels.data_ptr(), table.data_ptr(), None))
    e[1].record()
    nb64.zero_()
    e[2].record()
    # the bounce launch as the progressive frames issue it: rays formed in the kernel from the hit records, closest hits
    _lib.check(L.bmTraceBounceCamera(t.h, rays_dev.data_ptr(), n, W, f3(eye), f9(orient), 0.0, 0.0, 0xFFFFFFFF, 1, 0, hits1.data_ptr(), 0, 1, 1, hits2.data_ptr(), 0.0, 0.0, None, None, nb64.data_ptr(), None))
    e[3].record(); torch.cuda.synchronize()
    nb = int(nb64.item())
    print("rep %d: primary %.3f ms, bounce launch %.3f ms (%d rays, %.1f Mrays/s)" % (r, e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]), nb, nb / e[2].elapsed_time(e[3]) / 1e3), flush=True)
