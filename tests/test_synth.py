"""CPU tier: the synthetic workload generators: counter-based, reproducible, and identical between numpy and torch."""
import numpy as np
import torch


def test_soup_numpy_equals_torch_and_is_counter_based():
    from raytracercuda_b200 import synth
    P, I = synth.soup(5000)
    Pt, It = synth.soup_torch(5000, "cpu", chunk=1024)
    assert np.array_equal(P, Pt.numpy()) and np.array_equal(I.reshape(-1).view(np.int32), It.numpy())
    # counter based: any slice can be generated on its own
    P2, _ = synth.soup(1000, start=3000)
    assert np.array_equal(P2, P[9000:12000])
    c = P.reshape(-1, 3, 3).mean(1)
    assert np.abs(c).max() <= 28.02 and np.abs(P.reshape(-1, 3, 3) - c[:, None, :]).max() <= 0.0401
    assert abs(c.mean()) < 0.5 and c.std() > 15.0


def test_rays_numpy_equals_torch():
    from raytracercuda_b200 import synth
    R = synth.incoherent_rays(4000)
    Rt = synth.incoherent_rays_torch(4000, "cpu").numpy()
    assert np.array_equal(R[:, :3], Rt[:, :3]) and np.abs(R[:, 3:] - Rt[:, 3:]).max() < 1e-6
    assert np.abs(np.linalg.norm(R[:, 3:], axis=1) - 1).max() < 1e-5 and np.abs(R[:, :3]).max() <= 28.0
    assert abs(R[:, 5].mean()) < 0.05                                  # uniform on the sphere: z uniform in [-1, 1]


def test_icosphere_counts():
    from raytracercuda_b200 import synth
    for sub in (0, 1, 3):
        P, I, N = synth.displaced_icosphere(sub, seed=1)
        assert I.shape[0] == 20 * 4 ** sub and P.shape[0] == 10 * 4 ** sub + 2 and N.shape == P.shape
        r = np.linalg.norm(P, axis=1)
        assert 0.8 < r.min() and r.max() < 1.2
        # closed manifold: every edge is shared by exactly two triangles
        e = np.sort(np.concatenate([I[:, [0, 1]], I[:, [1, 2]], I[:, [2, 0]]]).astype(np.int64), axis=1)
        _, cnt = np.unique(e[:, 0] * (P.shape[0] + 1) + e[:, 1], return_counts=True)
        assert (cnt == 2).all()


def test_band_rows_partition_the_image():
    """strong scaling (bench.py --scaling strong): every image row is rendered by exactly one rank, in bands of 4 rows dealt round robin"""
    from raytracercuda_b200 import workloads
    for height, world in ((2160, 8), (1080, 8), (1080, 4), (720, 2), (37, 3), (4, 8)):
        parts = [workloads.band_rows(height, r, world) for r in range(world)]
        allrows = np.concatenate(parts)
        assert np.array_equal(np.sort(allrows), np.arange(height))
        sizes = [p.size for p in parts]
        assert max(sizes) - min(sizes) <= 4
        for r, p in enumerate(parts):
            assert ((p // 4) % world == r).all() and (np.diff(p) > 0).all()
