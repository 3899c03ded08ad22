import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import host, workloads
from raytracercuda_b200.tree import Tree
dev = torch.device("cuda:0")
wl = workloads.get("c2"); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); cam.setInitialRays(W, H, *wl.frustum)
rays = torch.from_numpy(cam.initialRays()).to(dev)
t = Tree(dev); t.add_mesh(wl.positions, wl.indices); t.build()
for y0 in range(180, 840, 60):
    sub = rays[y0 * W:(y0 + 60) * W]
    h, cc = t.trace_camera(sub, wl.eye, np.eye(3, dtype=np.float32), counters=True, width=W)
    torch.cuda.synchronize(); pk = int(cc.cpu()[3])
    print("rows %d: worst ray total %d = visits %d + clusters %d + tris %d" % (y0, pk >> 42, (pk >> 28) & 0x3FFF, (pk >> 14) & 0x3FFF, pk & 0x3FFF))
