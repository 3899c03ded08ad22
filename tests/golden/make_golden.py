"""Generates tests/golden/*.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Sources of truth:
  * meshes.npz      -- the reference's bundled assets (Content/f16.obj, suzanne.obj, bunny.zip) parsed by
                       raytracercuda_b200.meshio (positions / faces exactly as in the files).  The GPU box has no
                       /root/reference, so the parity tests read the meshes from here.
  * golden.json     -- produced by the UNMODIFIED reference compiled in its CUDA-0 host mode (oracle/_ref/libbeam_ref.so):
                       tree statistics, FNV-1a-64 fingerprints of the sorted (leaf code, triangle) list, hit counts and pixel
                       checksums for the reference viewer's default camera, primitive known answers, struct sizes, hash keys.
  * pixels_*.npz    -- the reference's own shaded pixels (200x200) for each mesh.
The oracle (oracle/oracle.cpp) and the CUDA path are both tested against these files.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import RefHashLib, RefLib, RefScene, fnv1a64_pairs  # noqa: E402
from raytracercuda_b200.meshio import load_obj  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CONTENT = "/root/reference/Content"


def main():
    R = RefLib()
    golden = {"source": "d1kkop/RaytracerCuda compiled with CUDA 0 (oracle/build_ref.sh), g++ -O2 -ffp-contract=off", "meshes": {}}
    arrays = {}
    for name, fn in (("f16", "f16.obj"), ("suzanne", "suzanne.obj"), ("bunny", "bunny.zip")):
        m = load_obj(os.path.join(CONTENT, fn))[0]
        arrays[name + "_pos"] = m.positions; arrays[name + "_idx"] = m.indices; arrays[name + "_nrm"] = m.normals
        s = RefScene(R); s.add_mesh(m.positions, m.indices, m.normals); s.build()
        st = s.stats(); code, mesh, tri = s.pairs()
        g = dict(st)
        g["num_tris"] = int(m.indices.shape[0]); g["num_verts"] = int(m.positions.shape[0])
        g["fnv1a64_pairs"] = "%016x" % fnv1a64_pairs(code, tri)
        g["crc32_codes"] = zlib.crc32(code.tobytes()); g["crc32_tris"] = zlib.crc32(tri.tobytes())
        g["trace"] = {}
        for (W, H) in ((200, 200), (500, 500) if name == "f16" else (320, 180)):
            err, px, _ = s.trace(W, H, [0, 0, -2.1], np.eye(3, dtype=np.float32))
            assert err == 0
            g["trace"]["%dx%d" % (W, H)] = {"hits": int((px != (255 << 8)).sum()), "crc32_pixels": zlib.crc32(px.tobytes())}
            if (W, H) == (200, 200):
                np.savez_compressed(os.path.join(OUT, "pixels_%s_200x200.npz" % name), pixels=px)
        golden["meshes"][name] = g
        s.close()
        print(name, g)
    # f16 split per material (what Assimp hands the reference: two meshes, aa.xml:56,58 grids 15 and 2 blocks)
    ms = load_obj(os.path.join(CONTENT, "f16.obj"), split_by_material=True)
    golden["f16_material_groups"] = [int(m.indices.shape[0]) for m in ms]
    s = RefScene(R)
    for m in ms:
        s.add_mesh(m.positions, m.indices, m.normals)
    s.build(); code, mesh, tri = s.pairs()
    golden["f16_two_meshes"] = {"pairs": int(code.size), "crc32_codes": zlib.crc32(code.tobytes()), "crc32_mesh": zlib.crc32(mesh.tobytes()), "crc32_tris": zlib.crc32(tri.tobytes())}
    err, px, _ = s.trace(200, 200, [0, 0, -2.1], np.eye(3, dtype=np.float32))
    golden["f16_two_meshes"]["crc32_pixels_200x200"] = zlib.crc32(px.tobytes())
    s.close()
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), **arrays)

    # primitive known answers (inputs from TestProgram/Program.cpp:81-90 and :156-158)
    f = lambda *a: np.array(a, np.float32)
    d = f(-1, -1, -1) / np.float32(np.sqrt(np.float32(3)))
    # normalize as glm does: v * (1/sqrt(dot))
    dd = f(-1, -1, -1); dot = np.float32(dd[0] * dd[0] + dd[1] * dd[1]) + np.float32(dd[2] * dd[2])
    d = (dd * (np.float32(1) / np.sqrt(dot))).astype(np.float32)
    inv = (np.float32(1) / d).astype(np.float32)
    import ctypes as C
    u, v = C.c_float(), C.c_float()
    t = R.lib.ref_tri_intersect(f(0, 0, -2.1), f(0, 0, 1), f(-1, -1, 1.56), f(0, 1, 1.56), f(1, -1, 1.56), C.byref(u), C.byref(v))
    golden["kat"] = {
        "box_ray_nozero_bits": int(np.float32(R.lib.ref_box_ray_nozero(f(0, 0, -3), f(1, 1, -2), f(0, 0, -2.3), inv)).view(np.uint32)),
        "box_ray_bits": int(np.float32(R.lib.ref_box_ray(f(0, 0, -3), f(1, 1, -2), f(0, 0, -2.3), inv)).view(np.uint32)),
        "tri_intersect": {"t_bits": int(np.float32(t).view(np.uint32)), "u": float(u.value), "v": float(v.value)},
    }
    sizes = np.zeros(10, np.uint32); R.lib.ref_struct_sizes(sizes)
    golden["struct_sizes"] = dict(zip(["bmFace", "bmFacePtr", "bmMaterial", "bmTreeNode", "bmFaceStore", "bmNodeStore", "bmFaceGroupStore",
                                       "StaticMeshData", "bmRayFaceHitResult", "bmStackNode"], [int(x) for x in sizes]))
    # random primitive vectors: SAT, ray/triangle, ray/box
    rng = np.random.default_rng(7)
    n = 4000
    cen = rng.uniform(-2, 2, (n, 3)).astype(np.float32); half = rng.uniform(0.01, 0.5, (n, 3)).astype(np.float32)
    tri = (cen[:, None, :] + rng.uniform(-0.8, 0.8, (n, 3, 3))).astype(np.float32)
    sat = np.array([R.lib.ref_tri_box_overlap(cen[i], half[i], np.ascontiguousarray(tri[i].reshape(-1))) for i in range(n)], np.uint8)
    org = rng.uniform(-3, 3, (n, 3)).astype(np.float32); dr = rng.normal(size=(n, 3)).astype(np.float32)
    dr[::17, 0] = 0.0      # axis-parallel directions exercise the inf / NaN paths
    tt = np.zeros(n, np.float32); uu = np.zeros(n, np.float32); vv = np.zeros(n, np.float32); bb = np.zeros(n, np.float32)
    for i in range(n):
        tt[i] = R.lib.ref_tri_intersect(org[i], dr[i], tri[i, 0].copy(), tri[i, 1].copy(), tri[i, 2].copy(), C.byref(u), C.byref(v)); uu[i] = u.value; vv[i] = v.value
        with np.errstate(divide="ignore"):
            iv = (np.float32(1) / dr[i]).astype(np.float32)
        bb[i] = R.lib.ref_box_ray((cen[i] - half[i]).astype(np.float32), (cen[i] + half[i]).astype(np.float32), org[i], iv)
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), cen=cen, half=half, tri=tri, sat=sat, org=org, dir=dr, t=tt, u=uu, v=vv, box=bb)
    # hash keys (Hash.cu:17-64)
    Hh = RefHashLib()
    ks = [0, 1, 2, 255, 256, 65535, 65536, 0xFFFFFFFF, 4294966296, 12345678] + [int(x) for x in rng.integers(0, 2**32, 200)]
    golden["hash"] = {"bmHash": [[k, int(Hh.lib.ref_hash(k))] for k in ks],
                      "bmHash3": [[1, 2, 3, int(Hh.lib.ref_hash3(1, 2, 3))], [0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD, int(Hh.lib.ref_hash3(0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD))]],
                      "bmMap": [[x, int(Hh.lib.ref_map(x))] for x in [0.0, 0.029, 0.03, -0.001, 1.0, -29.99, 29.99, 0.06, -0.03, 12.3456]],
                      "cell_size": int(Hh.lib.ref_cell_size()), "max_hash_elements": int(Hh.lib.ref_max_hash_elements())}
    pts = rng.uniform(-30, 30, (2000, 3)).astype(np.float32)
    hk = np.array([Hh.lib.ref_hash3(*[int(np.uint32(np.int32(Hh.lib.ref_map(float(c))))) for c in p]) for p in pts], np.uint32)
    cells = np.array([[Hh.lib.ref_map(float(c)) for c in p] for p in pts], np.int32)
    np.savez_compressed(os.path.join(OUT, "hashkeys.npz"), pts=pts, hash=hk, cells=cells)
    # camera direction table checksum (Camera.cpp:43-72)
    rays = np.zeros((200 * 200, 3), np.float32)
    assert R.lib.ref_initial_rays(200, 200, -1, 1, -1, 1, 1, rays) == 0
    golden["initial_rays_200x200_crc32"] = zlib.crc32(rays.tobytes())
    rays = np.zeros((1280 * 720, 3), np.float32)
    assert R.lib.ref_initial_rays(1280, 720, -1, 1, -1, 1, 1, rays) == 0
    golden["initial_rays_1280x720_crc32"] = zlib.crc32(rays.tobytes())
    json.dump(golden, open(os.path.join(OUT, "golden.json"), "w"), indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
