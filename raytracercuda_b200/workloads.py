"""Bench / test workloads = BASELINE.json `configs`, with the stand-ins SURVEY.md section 8(d) defines for the assets the
reference tree no longer ships (armadillo.obj, tyra.obj are missing large blobs)."""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_MESHES = os.path.join(ROOT, "tests", "golden", "meshes.npz")


@dataclass
class Workload:
    name: str                  # goes into the bench line's config.workload
    description: str
    positions: np.ndarray      # (V,3) float32
    indices: np.ndarray        # (F,3) uint32
    normals: np.ndarray        # (V,3) float32
    width: int
    height: int
    bounces: int               # diffuse bounce generations traced after the primary rays
    eye: tuple = synth.EYE
    frustum: tuple = synth.FRUSTUM
    samples: int = 1           # jittered samples per pixel (1 = the exact pixel-centre table, as the reference traces it)
    key: str = ""              # the name get() was called with

    @property
    def num_tris(self):
        return int(self.indices.shape[0])


def _bundled(name):
    z = np.load(GOLDEN_MESHES)
    return z[name + "_pos"], z[name + "_idx"], z[name + "_nrm"]


def get(name: str) -> Workload:
    wl = _get(name)
    wl.key = name
    return wl


def _get(name: str) -> Workload:
    if name == "c1":
        P, I, N = _bundled("bunny")
        return Workload("c1: Content/bunny.obj (69,630 tris) BVH build + 1 spp primary closest-hit at 1280x720", "configs[0]", P, I, N, 1280, 720, 0)
    if name == "c2":
        P, I, N = synth.displaced_icosphere(7, seed=1)
        return Workload("c2: armadillo STAND-IN (displaced icosphere, 327,680 tris, seed 1; Content/armadillo.obj is a missing blob) BVH build + "
                        "primary and 1-bounce diffuse rays at 1920x1080", "configs[1]", P, I, N, 1920, 1080, 1)
    if name == "c2small":        # same scene class at a size the CPU test tier can afford
        P, I, N = synth.displaced_icosphere(4, seed=1)
        return Workload("c2small: displaced icosphere 5,120 tris, 320x180 primary + 1 bounce", "test", P, I, N, 320, 180, 1)
    if name == "c3":
        P, I, N = _bundled("f16")
        return Workload("c3: Content/f16.obj (4,056 tris; the reference never reads the F16 textures) 64 spp jittered primary + 1 diffuse bounce at 1920x1080",
                        "configs[2]", P, I, N, 1920, 1080, 1, samples=64)
    if name == "c3small":
        P, I, N = _bundled("f16")
        return Workload("c3small: f16, 4 spp + 1 bounce at 320x180", "test", P, I, N, 320, 180, 1, samples=4)
    if name == "c4":
        # Radius 4 (eye moved back with it): at the reference viewer's unit scale 5.2 M triangles would put ~600 faces into every 0.03 leaf
        # cell, far beyond the reference's 256-face cap (BuildTree.cuh:17), where parity is not defined.  At this scale the
        # largest leaf holds 156 faces (overflow_leaves == 0; 7,991,992 pairs, 151,134 leaves -- checked against the oracle).
        P, I, N = synth.displaced_icosphere(9, seed=2, radius=4.0, amplitude=0.6)
        return Workload("c4: tyra STAND-IN (displaced icosphere radius 4, 5,242,880 tris, seed 2; Content/tyra.obj is a missing blob) 3840x2160 16 spp jittered primary + 1 diffuse bounce",
                        "configs[3]", P, I, N, 3840, 2160, 1, eye=(0.0, 0.0, -8.4), samples=16)
    if name == "c4small":
        P, I, N = synth.displaced_icosphere(5, seed=2, radius=4.0, amplitude=0.6)
        return Workload("c4small: displaced icosphere radius 4, 20,480 tris, 320x180 2 spp + 1 bounce", "test", P, I, N, 320, 180, 1, eye=(0.0, 0.0, -8.4), samples=2)
    raise ValueError(name)


def band_rows(height, rank, world, band=4):
    """Rows of a `height`-row image that rank `rank` of `world` renders in a strong-scaling frame: the bands of `band` consecutive rows
    rank, rank + world, ...  (band = the height of the trace kernels' 8x4 pixel tiles, so a rank's tiles are tiles of neighbouring pixels
    of the real image).  Every row belongs to exactly one rank; ranks may differ by one band."""
    return np.array([y for y in range(height) if (y // band) % world == rank], np.int64)


def interleaved_rows_frustum(frustum, height, rank, world):
    """Frustum (left,right,top,bottom,zoom) whose `height` rows are rows rank, rank+world, ... of the image that has
    height*world rows over the full frustum: ICamera::setInitialRays(W, height, *this) then yields exactly those rows
    (Camera.cpp:50-57: ry = top + dy/2 + k*dy with dy = (bottom-top)/height)."""
    left, right, top, bottom, zoom = frustum
    dy_full = (bottom - top) / (height * world)
    top_r = top + (rank + 0.5 - 0.5 * world) * dy_full
    return (left, right, top_r, top_r + height * world * dy_full, zoom)
