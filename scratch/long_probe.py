import sys, os, numpy as np, torch, ctypes as C
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import _lib, host, workloads
from raytracercuda_b200.tree import Tree, hits_to_numpy
L = _lib.load(); dev = torch.device("cuda:0")
wl = workloads.get(sys.argv[1] if len(sys.argv) > 1 else "c2"); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); cam.setInitialRays(W, H, *wl.frustum)
rays = torch.from_numpy(cam.initialRays()).to(dev)
t = Tree(dev); t.add_mesh(wl.positions, wl.indices); t.build()
hits, cnt = t.trace_camera(rays, wl.eye, np.eye(3, dtype=np.float32), counters=True, width=W)
torch.cuda.synchronize()
c = cnt.cpu().numpy(); print("totals per ray: visits %.2f tris %.2f clusters %.2f ; longest single ray: %d steps" % (c[0] / n, c[1] / n, c[2] / n, c[3]))
# row-by-row cost: trace each row band of 8 rows separately with counters to see where the long rays live
ev = lambda: torch.cuda.Event(enable_timing=True)
worst = []
for y0 in range(0, H, 60):
    sub = rays[y0 * W:(y0 + 60) * W]
    h, cc = t.trace_camera(sub, wl.eye, np.eye(3, dtype=np.float32), counters=True, width=W)
    e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record()
    t.trace_camera(sub, wl.eye, np.eye(3, dtype=np.float32), width=W)
    e1.record(); torch.cuda.synchronize()
    cc = cc.cpu().numpy(); nn = sub.shape[0]
    print("rows %4d-%4d: %.3f ms  steps/ray %.1f  longest %d  hits %d" % (y0, y0 + 59, e0.elapsed_time(e1), (cc[0] + cc[1] + cc[2]) / nn, cc[3], int((hits_to_numpy(h)["tri"] != 0xFFFFFFFF).sum())))
