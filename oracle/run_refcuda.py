"""TEST / BASELINE INFRASTRUCTURE ONLY.  Runs the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref/libbeam_refcuda.so)
in a process of their own and prints one JSON line.  A separate process because the reference's build path is racy by design
(BuildTree.cu:13-18,43-50: allocations lost to a failed atomicCAS are leaked) and can overrun its fixed pools under B200-scale
parallelism, which ends in a sticky CUDA error / exit(1) (CudaComon.cuh:59-68).

    python oracle/run_refcuda.py <mesh name in tests/golden/meshes.npz | c2> <W> <H> [repeat]
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name, W, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    repeat = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    from oracle.bindings import RefGpuScene
    if name in ("f16", "suzanne", "bunny"):
        z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
        P, I, N = z[name + "_pos"], z[name + "_idx"], z[name + "_nrm"]
    else:
        from raytracercuda_b200 import workloads
        wl = workloads.get(name)
        P, I, N = wl.positions, wl.indices, wl.normals
    s = RefGpuScene("refcuda")
    s.add_mesh(P, I, N)
    s.build()                                   # warm-up (first launch, allocations)
    tb = min(s.build() for _ in range(3))
    s.camera(W, H)
    s.trace([0, 0, -2.1], np.eye(3, dtype=np.float32), 1)
    tt = s.trace([0, 0, -2.1], np.eye(3, dtype=np.float32), repeat)
    px = s.pixels()
    out = {"mesh": name, "triangles": int(I.shape[0]), "width": W, "height": H, "build_ms": tb * 1e3, "trace_ms": tt * 1e3,
           "build_mtris_per_s": I.shape[0] / tb / 1e6, "primary_mrays_per_s": W * H / tt / 1e6,
           "hits": int((px != 255 << 8).sum()), "crc32_pixels": zlib.crc32(px.tobytes())}
    gp = os.path.join(ROOT, "tests", "golden", "pixels_%s_%dx%d.npz" % (name, W, H))
    if os.path.exists(gp):
        want = np.load(gp)["pixels"]
        out["pixel_equality_vs_reference_cpu_mode"] = float((px == want).mean())
        out["hit_miss_agreement"] = float(((px != 255 << 8) == (want != 255 << 8)).mean())
    print(json.dumps(out), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
