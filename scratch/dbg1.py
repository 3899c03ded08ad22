import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from oracle.bindings import OracleLib, OracleScene
from raytracercuda_b200.tree import Tree, hits_to_numpy, make_mesh_table
z = np.load('tests/golden/meshes.npz'); P, I, N = z['bunny_pos'], z['bunny_idx'], z['bunny_nrm']
want = np.load('tests/golden/pixels_bunny_200x200.npz')['pixels'].reshape(-1)
O = OracleLib(); s = OracleScene(O); s.add_mesh(P, I, N); s.build()
rays = O.initial_rays(200, 200)
r = s.trace_camera(rays, [0,0,-2.1], np.eye(3,dtype=np.float32), 0, 8)
print('oracle==golden', np.array_equal(r['pixels'], want))
t = Tree('cuda:0'); t.add_mesh(P, I); t.build()
Pd, Id = t._meshes[0]
table = make_mesh_table([(Pd, Id, torch.from_numpy(N).cuda())], 'cuda:0')
px = torch.empty(40000, dtype=torch.int32, device='cuda')
hits,_ = t.trace_camera(rays, [0,0,-2.1], np.eye(3,dtype=np.float32), pixels=px, mesh_table=table)
h = hits_to_numpy(hits); g = px.cpu().numpy().view(np.uint32)
bad = np.flatnonzero(g != want)
print('mismatch pixels', bad.size, bad[:10])
for i in bad[:10]:
    print(i, hex(g[i]), hex(want[i]), 'gpu tri', h['tri'][i], 't', h['t'][i], 'u', h['u'][i], 'v', h['v'][i], '| orc tri', r['tri'][i], 't', r['t'][i], r['u'][i], r['v'][i], 'tie', r['tie'][i])
