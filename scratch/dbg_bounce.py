import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from raytracercuda_b200 import _lib, host, workloads
from raytracercuda_b200.tree import Tree
dev = torch.device("cuda", 0); L = _lib.load(); wl = workloads.get(sys.argv[1] if len(sys.argv) > 1 else "c2"); W, H = wl.width, wl.height; n = W * H
cam = host.ICamera.create(); assert cam.setInitialRays(W, H, *wl.frustum) == 0
table = torch.from_numpy(cam.initialRays()).to(dev)
t = Tree(dev); t.add_mesh(wl.positions, wl.indices); t.build(); print(t.stats(), flush=True)
eye = (C.c_float * 3)(*[float(x) for x in wl.eye]); ori = (C.c_float * 9)(*np.eye(3, dtype=np.float32).reshape(-1))
h1 = torch.empty((n, 4), dtype=torch.float32, device=dev); h2 = torch.empty((n, 4), dtype=torch.float32, device=dev)
r2 = torch.empty((n, 6), dtype=torch.float32, device=dev); src = torch.empty(n, dtype=torch.int32, device=dev); cnt2 = torch.zeros(1, dtype=torch.int32, device=dev)
nb = torch.zeros(1, dtype=torch.int64, device=dev); cnt = torch.zeros(4, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
_lib.check(L.bmTraceCamera(t.h, table.data_ptr(), n, W, eye, ori, 0, h1.data_ptr(), None, None, None))
_lib.check(L.bmBounceRaysCamera(t.h, table.data_ptr(), eye, ori, h1.data_ptr(), n, 0, 1, 1, r2.data_ptr(), src.data_ptr(), cnt2.data_ptr()))
torch.cuda.synchronize(); print("list count", int(cnt2.item()), "hits", int((h1[:, 3].view(torch.int32) != -1).sum()))
for counters in (None, cnt):
    nb.zero_(); cnt.zero_()
    _lib.check(L.bmTraceBounceCamera(t.h, table.data_ptr(), n, W, eye, ori, 0.0, 0.0, 0xFFFFFFFF, 1, 0, h1.data_ptr(), 0, 1, 1, h2.data_ptr(), 0.0, 0.0, None, None, nb.data_ptr(), None if counters is None else counters.data_ptr()))
    torch.cuda.synchronize(); print("bounce launch count", int(nb.item()), "counters", cnt.tolist())
for rep in range(5):
    tt = []
    for which in range(3):
        flush.zero_(); e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record()
        if which == 0: _lib.check(L.bmTraceCamera(t.h, table.data_ptr(), n, W, eye, ori, 0, h1.data_ptr(), None, None, None))
        elif which == 1: _lib.check(L.bmTraceBounceCamera(t.h, table.data_ptr(), n, W, eye, ori, 0.0, 0.0, 0xFFFFFFFF, 1, 0, h1.data_ptr(), 0, 1, 1, h2.data_ptr(), 0.0, 0.0, None, None, nb.data_ptr(), None))
        else: _lib.check(L.bmTraceRaysIndirect(t.h, r2.data_ptr(), cnt2.data_ptr(), n, 1, h2.data_ptr(), None))
        e1.record(); torch.cuda.synchronize(); tt.append(e0.elapsed_time(e1))
    print("rep %d: primary %.3f ms, bounce launch %.3f ms, bounce list trace %.3f ms" % (rep, *tt), flush=True)
