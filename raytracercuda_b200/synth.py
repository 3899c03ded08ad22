"""Synthetic inputs defined by SURVEY.md section 8(d): the stand-in meshes for the absent armadillo / tyra assets, the
counter-based triangle soup and incoherent ray batches of config 5, and the camera defaults of the reference viewer."""
from __future__ import annotations

import numpy as np

EYE = (0.0, 0.0, -2.1)                 # TestProgram/Program.cpp:106
FRUSTUM = (-1.0, 1.0, -1.0, 1.0, 1.0)  # setInitialRays(w, h, -1, 1, -1, 1, 1), Program.cpp:189
IDENTITY = np.eye(3, dtype=np.float32)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform(seed: int, k: np.ndarray, lane: int) -> np.ndarray:
    """counter-based uniform [0,1) float64 from splitmix64(seed, k, lane)"""
    with np.errstate(over="ignore"):
        x = _splitmix64(np.uint64(seed) ^ _splitmix64(k.astype(np.uint64) * np.uint64(16) + np.uint64(lane)))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def soup(num_tris: int, seed: int = 12345, extent: float = 28.0, edge: float = 0.02, start: int = 0):
    """Config-5 soup: triangle k has a centre uniform in [-extent,extent]^3 and vertices = centre + uniform[-edge,edge]^3.
    Returns (positions (3n,3) float32, indices (n,3) uint32)."""
    k = np.arange(start, start + num_tris, dtype=np.uint64)
    c = np.stack([(_uniform(seed, k, a) * 2 - 1) * extent for a in range(3)], axis=1)
    P = np.empty((num_tris, 3, 3), dtype=np.float32)
    for v in range(3):
        for a in range(3):
            P[:, v, a] = (c[:, a] + (_uniform(seed, k, 3 + 3 * v + a) * 2 - 1) * edge).astype(np.float32)
    I = np.arange(3 * num_tris, dtype=np.uint32).reshape(-1, 3)
    return P.reshape(-1, 3), I


def incoherent_rays(num: int, seed: int = 54321, extent: float = 28.0, start: int = 0) -> np.ndarray:
    """Config-5 rays: origin uniform in [-extent,extent]^3, direction uniform on the sphere. (n,6) float32 bmRay records."""
    k = np.arange(start, start + num, dtype=np.uint64)
    o = np.stack([(_uniform(seed, k, a) * 2 - 1) * extent for a in range(3)], axis=1)
    z = _uniform(seed, k, 3) * 2 - 1
    phi = _uniform(seed, k, 4) * 2 * np.pi
    r = np.sqrt(np.maximum(0.0, 1 - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    return np.concatenate([o, d], axis=1).astype(np.float32)


def _value_noise(p: np.ndarray, seed: int) -> np.ndarray:
    """smooth 3-D value noise in [-1,1] (trilinear blend of hashed lattice values, 3 octaves)"""
    def lattice(ix, iy, iz, s):
        with np.errstate(over="ignore"):
            h = _splitmix64((ix.astype(np.int64).astype(np.uint64) * np.uint64(73856093)) ^ (iy.astype(np.int64).astype(np.uint64) * np.uint64(19349663))
                            ^ (iz.astype(np.int64).astype(np.uint64) * np.uint64(83492791)) ^ np.uint64(s))
        return (h >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
    out = np.zeros(p.shape[0])
    amp, freq = 1.0, 2.0
    for o in range(3):
        q = p * freq
        i = np.floor(q); f = q - i
        f = f * f * (3 - 2 * f)
        ix, iy, iz = i[:, 0], i[:, 1], i[:, 2]
        acc = 0
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                    acc = acc + w * lattice(ix + dx, iy + dy, iz + dz, seed * 1315423911 + o)
        out += amp * acc
        amp *= 0.5; freq *= 2.0
    return out / 1.75


def displaced_icosphere(subdiv: int, seed: int = 1, radius: float = 1.0, amplitude: float = 0.15):
    """Closed icosphere with 20*4^subdiv triangles, radius displaced by value noise: the STAND-IN for the reference's missing
    Content/armadillo.obj (subdiv 7 = 327,680 triangles, seed 1) and Content/tyra.obj (subdiv 9 = 5,242,880 triangles, seed 2).
    Returns (positions (V,3) float32, indices (F,3) uint32, normals (V,3) float32)."""
    t = (1.0 + 5 ** 0.5) / 2.0
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdiv):
        nV = V.shape[0]
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * np.int64(nV) + es[:, 1]
        uk, inv = np.unique(key, return_inverse=True)
        a = uk // nV; b = uk % nV
        mid = V[a] + V[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        V = np.concatenate([V, mid], axis=0)
        m = nV + inv.reshape(3, -1).T          # per face: midpoints of edges (01, 12, 20)
        F = np.concatenate([np.stack([F[:, 0], m[:, 0], m[:, 2]], 1), np.stack([F[:, 1], m[:, 1], m[:, 0]], 1),
                            np.stack([F[:, 2], m[:, 2], m[:, 1]], 1), np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)], axis=0)
    r = radius + amplitude * _value_noise(V, seed)
    P = (V * r[:, None]).astype(np.float32)
    from .meshio import compute_vertex_normals
    I = F.astype(np.uint32)
    return P, I, compute_vertex_normals(P, I)


def orient_from_pan_pitch(pan: float, pitch: float) -> np.ndarray:
    """column-major 3x3 (glm::mat3 layout) for a yaw/pitch camera: returned flat as 9 floats m[3*col+row]"""
    cy, sy, cp, sp = np.cos(pan), np.sin(pan), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    M = (Ry @ Rx).astype(np.float32)
    return np.ascontiguousarray(M.T).reshape(-1)      # row-major of the transpose == column-major of M


# ---- the same generators on a torch device (config 5 at full size: 100 M triangles are generated where they are used) ----------------
def _t_lsr(x, s):
    import torch
    return (x >> s) & ((1 << (64 - s)) - 1)                       # logical shift right on int64


def _t_splitmix64(x):
    import torch
    x = x + torch.tensor(0x9E3779B97F4A7C15 - (1 << 64), dtype=torch.int64, device=x.device)
    z = x
    z = (z ^ _t_lsr(z, 30)) * torch.tensor(0xBF58476D1CE4E5B9 - (1 << 64), dtype=torch.int64, device=x.device)
    z = (z ^ _t_lsr(z, 27)) * torch.tensor(0x94D049BB133111EB - (1 << 64), dtype=torch.int64, device=x.device)
    return z ^ _t_lsr(z, 31)


def _t_uniform(seed, k, lane):
    import torch
    x = _t_splitmix64(torch.tensor(seed, dtype=torch.int64, device=k.device) ^ _t_splitmix64(k * 16 + lane))
    return _t_lsr(x, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def soup_torch(num_tris: int, device, seed: int = 12345, extent: float = 28.0, edge: float = 0.02, start: int = 0, chunk: int = 1 << 22):
    """soup() evaluated with torch on `device`, bit-identical to the numpy version; returns (positions (3n,3) float32, indices (3n,) int32)"""
    import torch
    P = torch.empty((num_tris, 3, 3), dtype=torch.float32, device=device)
    for a0 in range(0, num_tris, chunk):
        a1 = min(num_tris, a0 + chunk)
        k = torch.arange(start + a0, start + a1, dtype=torch.int64, device=device)
        c = [(_t_uniform(seed, k, a) * 2 - 1) * extent for a in range(3)]
        for v in range(3):
            for a in range(3):
                P[a0:a1, v, a] = (c[a] + (_t_uniform(seed, k, 3 + 3 * v + a) * 2 - 1) * edge).to(torch.float32)
    I = torch.arange(3 * num_tris, dtype=torch.int32, device=device)
    return P.reshape(-1, 3), I


def incoherent_rays_torch(num: int, device, seed: int = 54321, extent: float = 28.0, start: int = 0, out=None):
    """incoherent_rays() evaluated with torch on `device` (same counter-based construction; float64 trig, float32 result)"""
    import torch
    k = torch.arange(start, start + num, dtype=torch.int64, device=device)
    R = out if out is not None else torch.empty((num, 6), dtype=torch.float32, device=device)
    for a in range(3):
        R[:, a] = ((_t_uniform(seed, k, a) * 2 - 1) * extent).to(torch.float32)
    z = _t_uniform(seed, k, 3) * 2 - 1
    phi = _t_uniform(seed, k, 4) * (2 * np.pi)
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    R[:, 3] = (r * torch.cos(phi)).to(torch.float32); R[:, 4] = (r * torch.sin(phi)).to(torch.float32); R[:, 5] = z.to(torch.float32)
    return R
