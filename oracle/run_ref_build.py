"""TEST / BASELINE INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference in its CUDA-0 host mode (oracle/_ref/libbeam_ref.so, built by
oracle/build_ref.sh) on one of the bench workloads in a process of its own and prints one JSON line: build seconds, trace seconds and
the pixels (written to an .npy file).  A separate process because the reference's fixed-size pools (BuildTree.cuh:12-17) overflow on
scenes beyond what it was written for, which ends in a segmentation fault rather than an error code.

    python oracle/run_ref_build.py <workload> <W> <H> <pixels.npy>
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name, W, H, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    from oracle.bindings import RefLib, RefScene
    from raytracercuda_b200 import workloads
    wl = workloads.get(name)
    s = RefScene(RefLib())
    s.add_mesh(wl.positions, wl.indices, wl.normals)
    tb = s.build()
    err, px, sec = s.trace(W, H, wl.eye, np.eye(3, dtype=np.float32), wl.frustum)
    s.close()
    if not err:
        np.save(out, np.asarray(px, np.uint32).reshape(-1))
    print(json.dumps({"error": int(err), "build_seconds": tb, "trace_seconds": sec, "triangles": int(wl.num_tris)}), flush=True)


if __name__ == "__main__":
    main()
