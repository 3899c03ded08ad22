"""scratch: build a soup of --mtris million triangles and trace --mrays million incoherent rays once per mode (for ncu captures)"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from raytracercuda_b200 import synth, _lib
from raytracercuda_b200.tree import Tree
ap = argparse.ArgumentParser(); ap.add_argument("--mtris", type=float, default=20); ap.add_argument("--mrays", type=float, default=8)
ap.add_argument("--modes", default="0"); ap.add_argument("--bricks", type=int, default=-1); ap.add_argument("--bits", type=int, default=0); ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = "cuda:0"
T = int(a.mtris * 1e6); n = int(a.mrays * 1e6)
P, I = synth.soup_torch(T, dev)
t = Tree(dev); t.add_mesh(P, I); t.build(); torch.cuda.synchronize()
print(t.stats(), flush=True)
t.set_bricks(a.bricks, a.bits)
R = synth.incoherent_rays_torch(n, dev)
hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
for mode in [int(x) for x in a.modes.split(",")]:
    for r in range(a.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cnt = torch.zeros(4, dtype=torch.int64, device=dev) if r == 0 else None
        e0.record(); t.trace_rays(R, mode, hits_out=hits) if cnt is None else _lib.check(t.L.bmTraceRays(t.h, R.data_ptr(), n, mode, hits.data_ptr(), cnt.data_ptr())); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        extra = "" if cnt is None else " counters(visits,tris,clusters,longest)=%s per ray %.1f" % (cnt.tolist(), cnt[0].item() / n)
        print("mode %d rep %d: %.2f ms = %.1f Mrays/s hit %.3f%s" % (mode, r, ms, n / ms / 1e3, float((hits[:, 3].view(torch.int32) != -1).float().mean()), extra), flush=True)
