import sys, os, numpy as np, torch, time
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import workloads
from raytracercuda_b200.tree import Tree
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
if name.startswith("soup"):
    from raytracercuda_b200 import synth
    P, I = synth.soup(int(name[4:]) * 1000000)
else:
    wl = workloads.get(name); P, I = wl.positions, wl.indices
dev = "cuda:0"
Pd = torch.from_numpy(P).to(dev); Id = torch.from_numpy(I.view(np.int32)).to(dev).reshape(-1)
t = Tree(dev)
for i in range(int(os.environ.get("NB", "4"))):
    t.reset(); t.add_mesh(Pd, Id)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t.build()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(name, t.stats(), "build %.3f ms -> %.1f Mtris/s" % (dt * 1e3, I.shape[0] / dt / 1e6))
