"""primary-ray launch time with temporal tile scheduling while the camera MOVES between frames (orbit step in degrees per frame)"""
import sys, os, math, ctypes as C, numpy as np, torch
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for step_deg in (0.0, 0.25, 1.0, 5.0, 20.0):
    ms = []
    for f in range(24):
        a = math.radians(step_deg * f)
        eye = (C.c_float * 3)(2.1 * math.sin(a), 0.0, -2.1 * math.cos(a))
        c, s = math.cos(a), math.sin(a)
        # camera looks at the origin: rotate the view about y by -a (column-major mat3)
        orient = (C.c_float * 9)(c, 0, s, 0, 1, 0, -s, 0, c)
        flush.zero_(); e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record()
        _lib.check(L.bmTraceCamera(tree.h, rays.data_ptr(), n, W, eye, orient, 0, hits.data_ptr(), pix.data_ptr(), table.data_ptr(), None), "trace")
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    hit = int((hits[:, 3].view(torch.int32) != -1).sum())
    print("orbit %5.2f deg/frame: first frame %.3f ms, mean of frames 4.. %.3f ms, hits %d" % (step_deg, ms[0], sum(ms[4:]) / len(ms[4:]), hit), flush=True)
    # forget the history between settings
    tree.reset(); tree.add_mesh(Pd, Id); tree.build()
