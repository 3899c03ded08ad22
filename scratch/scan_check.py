# exercises the scan through the builds it drives: soup builds of several sizes vs the oracle pair counts are in the test-suite;
# here: time the c2 build and a 20 M soup build
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import workloads, synth
from raytracercuda_b200.tree import Tree
dev = "cuda:0"
for name in ("c2", "soup2", "soup20"):
    if name.startswith("soup"):
        P, I = synth.soup(int(name[4:]) * 1000000)
    else:
        wl = workloads.get(name); P, I = wl.positions, wl.indices
    Pd = torch.from_numpy(P).to(dev); Id = torch.from_numpy(I.view(np.int32)).to(dev).reshape(-1)
    t = Tree(dev)
    for i in range(5):
        t.reset(); t.add_mesh(Pd, Id)
        torch.cuda.synchronize(); t0 = time.perf_counter(); t.build(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, t.stats(), "build %.3f ms -> %.1f Mtris/s" % (dt * 1e3, I.shape[0] / dt / 1e6), flush=True)
    t.close()
