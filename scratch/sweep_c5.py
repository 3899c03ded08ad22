"""HBM-bound sweep (BASELINE config 5 scaled to one GPU's time budget): counter-based triangle soup, incoherent rays.
python scratch/sweep_c5.py <Mtris> <Mrays>"""
import sys, os, json, time, numpy as np, torch, ctypes as C
sys.path.insert(0, os.getcwd())
from raytracercuda_b200 import synth, _lib
from raytracercuda_b200.tree import Tree
mt = float(sys.argv[1]); mr = float(sys.argv[2])
dev = "cuda:0"
t0 = time.time(); P, I = synth.soup(int(mt * 1e6)); tg = time.time() - t0
Pd = torch.from_numpy(P).to(dev); Id = torch.from_numpy(I.view(np.int32)).to(dev).reshape(-1)
t = Tree(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
bms = []
for i in range(3):
    t.reset(); t.add_mesh(Pd, Id)
    e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record(); t.build(); e1.record(); torch.cuda.synchronize(); bms.append(e0.elapsed_time(e1))
st = t.stats()
n = int(mr * 1e6)
g = torch.Generator(device=dev); g.manual_seed(54321)
R = torch.empty((n, 6), dtype=torch.float32, device=dev)
R[:, :3] = (torch.rand((n, 3), generator=g, device=dev) * 2 - 1) * 28.0
z = torch.rand(n, generator=g, device=dev) * 2 - 1; phi = torch.rand(n, generator=g, device=dev) * (2 * np.pi); r = torch.sqrt(torch.clamp(1 - z * z, min=0))
R[:, 3] = r * torch.cos(phi); R[:, 4] = r * torch.sin(phi); R[:, 5] = z
del z, phi, r
hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
L = _lib.load()
out = {"triangles": st["triangles"], "pairs": st["pairs"], "leaves": st["leaves"], "build_ms": min(bms), "build_mtris_per_s": st["triangles"] / min(bms) / 1e3, "rays": n}
for mode, name in ((0, "reference"), (1, "closest")):
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.check(L.bmTraceRays(t.h, R.data_ptr(), n, mode, hits.data_ptr(), cnt.data_ptr()), "trace")
    torch.cuda.synchronize()
    V, K, Cc, LG = [int(x) for x in cnt.cpu()]
    ms = []
    for i in range(3):
        e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record()
        _lib.check(L.bmTraceRays(t.h, R.data_ptr(), n, mode, hits.data_ptr(), None), "trace")
        e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    hit = float((hits[:, 3].view(torch.int32) != -1).float().mean())
    by = n * (24 + 16) + 96 * V + 32 * Cc + 48 * K
    out[name] = {"ms": min(ms), "mrays_per_s": n / min(ms) / 1e3, "hit_rate": hit, "node_visits_per_ray": V / n, "tri_tests_per_ray": K / n, "cluster_boxes_per_ray": Cc / n,
                 "bytes_per_ray": by / n, "achieved_gbs": by / min(ms) / 1e6, "frac_of_measured_hbm_6555.8": by / min(ms) / 1e6 / 6555.8}
print(json.dumps(out))
