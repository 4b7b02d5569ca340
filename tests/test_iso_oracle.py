"""oracle/iso_oracle.py (the numpy restatement of the device isosurface) against analytic surfaces: every vertex on the level
set of the sampled function along its lattice edge, a closed oriented manifold for a sphere inside the volume, its area,
NaN samples cutting a hole instead of producing geometry."""
import numpy as np

from oracle import iso_oracle as io


def sphere_volume(n=17, r=0.62):
    ax = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(ax, ax * 0.9, ax * 1.1, indexing="ij")
    return r - np.sqrt(x * x + y * y + z * z), (ax[1] - ax[0], 0.9 * (ax[1] - ax[0]), 1.1 * (ax[1] - ax[0]))


def edge_use(faces):
    from collections import Counter

    cnt = Counter()
    for a, b, c in faces:
        for e in ((a, b), (b, c), (c, a)):
            cnt[e] += 1
    return cnt


def test_sphere_is_a_closed_oriented_manifold_with_the_right_area():
    vol, sp = sphere_volume()
    verts, faces, normals = io.marching_tetrahedra(vol, 0.0, sp)
    assert len(verts) > 300 and len(faces) > 600
    cnt = edge_use(faces)
    # every directed edge once, and its reverse once: closed, consistently oriented
    assert set(cnt.values()) == {1} and all((b, a) in cnt for (a, b) in cnt)
    assert len(verts) - len(cnt) // 2 + len(faces) == 2  # Euler characteristic of a sphere
    p = verts - np.array(vol.shape) // 2 * np.array(sp)
    tri = p[faces]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
    assert abs(area - 4 * np.pi * 0.62**2) / (4 * np.pi * 0.62**2) < 0.03
    # inside (>= level) is the ball: triangle normals and vertex normals point outwards (towards decreasing values)
    cen = tri.mean(axis=1)
    assert np.all(np.einsum("ij,ij->i", nrm, cen) > 0)
    assert np.all(np.einsum("ij,ij->i", normals, p) > 0)
    np.testing.assert_allclose(np.linalg.norm(normals, axis=1), 1.0, atol=1e-12)
    # the vertices sit where the linear interpolant along their lattice edge reaches the level: within h^2 of the sphere
    assert np.abs(np.linalg.norm(p, axis=1) - 0.62).max() < 0.5 * max(sp) ** 2 / 0.62 + 1e-12


def test_nan_samples_cut_a_hole():
    vol, sp = sphere_volume(13)
    vol[8:, :, :] = np.nan
    verts, faces, _ = io.marching_tetrahedra(vol, 0.0, sp)
    assert len(faces) > 0 and np.isfinite(verts).all()
    assert verts[:, 0].max() <= 7 * sp[0] + 1e-12  # nothing beyond the last finite slab
    cnt = edge_use(faces)
    assert any((b, a) not in cnt for (a, b) in cnt)  # an open boundary where the volume ends


def test_a_plane_between_the_samples():
    vol = np.fromfunction(lambda i, j, k: 0.3 * i + 0.2 * j - 0.5 * k + 0.05, (4, 5, 6))
    verts, faces, normals = io.marching_tetrahedra(vol, 0.0, (1.0, 2.0, 0.5))
    f = 0.3 * verts[:, 0] + 0.2 * verts[:, 1] / 2.0 - 0.5 * verts[:, 2] / 0.5 + 0.05
    np.testing.assert_allclose(f, 0.0, atol=1e-13)  # a linear field: the piecewise-linear interpolant is exact
    g = -np.array([0.3, 0.1, -1.0])
    np.testing.assert_allclose(normals, np.tile(g / np.linalg.norm(g), (len(verts), 1)), atol=1e-12)
