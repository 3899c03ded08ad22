"""Generates tests/golden/f16_materials.npz from the reference's bundled Content/f16.obj + f16.mtl + F16s.bmp / F16t.bmp
(run in the container that has /root/reference; the GPU box only reads the fixture):
  uv            (V,2) float32  `vt` coordinates per position (first use wins), from the library's own loader (csrc/objload.cpp)
  idx_F16s/_F16t (F,3) uint32  the two `usemtl` groups = the two meshes Assimp hands the reference (aa.xml:56,58)
  tex_F16s/_F16t (h,w) uint32  the map_Kd images decimated to <= 64 texels per side (0x00RRGGBB, row 0 = top)
Only asset data derived from the reference's content files; no reference source."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from raytracercuda_b200 import meshio  # noqa: E402

CONTENT = os.environ.get("REF_CONTENT", "/root/reference/Content")


def main():
    ms = meshio.load_obj_native(os.path.join(CONTENT, "f16.obj"), split_by_material=True)
    out = {"uv": ms[0].uvs}
    for m in ms:
        out["idx_" + m.name] = m.indices
        tex = meshio.load_bmp_native(m.texture)
        sy, sx = max(1, tex.shape[0] // 64), max(1, tex.shape[1] // 64)
        out["tex_" + m.name] = np.ascontiguousarray(tex[::sy, ::sx])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "f16_materials.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
