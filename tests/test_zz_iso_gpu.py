"""lsb200_isosurface (csrc/iso.cuh) through the C-ABI: entry for entry against oracle/iso_oracle.py on small volumes, analytic
properties at a size the oracle's Python loops do not reach, and BaseFeature.surfaces of the unmodified reference through the
plug-in (the call chain of model['strati'].surfaces(value): regular grid -> evaluate_value -> marching_cubes -> Surface)."""
import os
import sys

import numpy as np
import pytest

from oracle import iso_oracle as io

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module", autouse=True)
def _built():
    from loopstructural_b200 import build

    build.build()


def field(shape, seed):
    rng = np.random.default_rng(seed)
    ax = [np.linspace(-1, 1, n) for n in shape]
    x, y, z = np.meshgrid(*ax, indexing="ij")
    return 0.6 - np.sqrt(x * x + 0.8 * y * y + 1.2 * z * z) + 0.15 * np.sin(3 * x + rng.uniform(0, 3)) * np.cos(2 * y)


@pytest.mark.parametrize("shape,level,spacing,holes", [((9, 8, 7), 0.0, (1.0, 1.0, 1.0), 0), ((6, 11, 5), 0.1, (0.5, 2.0, 1.5), 0),
                                                        ((8, 8, 9), -0.05, (1.0, 0.7, 1.3), 12), ((2, 2, 2), 0.0, (1.0, 1.0, 1.0), 0)])
def test_device_isosurface_equals_the_oracle(shape, level, spacing, holes):
    from loopstructural_b200.isosurface import marching_tetrahedra

    vol = field(shape, sum(shape))
    if shape == (2, 2, 2):
        vol = np.arange(8, dtype=float).reshape(shape) - 3.4
    if holes:
        idx = np.random.default_rng(1).integers(0, vol.size, holes)
        vol.reshape(-1)[idx] = np.nan
    v_ref, f_ref, n_ref = io.marching_tetrahedra(vol, level, spacing)
    verts, faces, normals, values = marching_tetrahedra(vol, level, spacing=spacing)
    assert verts.shape == v_ref.shape and faces.shape == f_ref.shape and faces.dtype == np.int32
    np.testing.assert_allclose(verts, v_ref, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(faces, f_ref)
    np.testing.assert_allclose(normals, n_ref, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(values, np.full(len(verts), level))


def test_large_sphere_is_closed_and_has_the_right_area():
    from loopstructural_b200.isosurface import marching_tetrahedra

    n, r = 160, 0.7
    ax = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    vol = r - np.sqrt(x * x + y * y + z * z)
    h = ax[1] - ax[0]
    verts, faces, normals, _ = marching_tetrahedra(vol, 0.0, spacing=(h, h, h))
    p = verts - 1.0
    assert np.abs(np.linalg.norm(p, axis=1) - r).max() < h * h
    tri = p[faces]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
    assert abs(area / (4 * np.pi * r * r) - 1) < 2e-3
    assert np.all(np.einsum("ij,ij->i", nrm, tri.mean(axis=1)) > 0) and np.all(np.einsum("ij,ij->i", normals, p) > 0)
    # closed: every directed edge has its reverse exactly once
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * (len(verts) + 1) + e[:, 1]
    rev = e[:, 1] * (len(verts) + 1) + e[:, 0]
    assert len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rev))
    assert len(verts) - len(key) // 2 + len(faces) == 2


def test_errors_are_the_callers():
    from loopstructural_b200.isosurface import marching_tetrahedra

    vol = field((5, 5, 5), 0)
    with pytest.raises(ValueError):
        marching_tetrahedra(vol, 99.0)
    with pytest.raises(ValueError):
        marching_tetrahedra(vol[0], 0.0)


def test_surfaces_of_a_model_feature_through_the_plug_in():
    """model['strati'].surfaces(value) with the plug-in installed: LoopIsosurfacer.fit evaluates the feature on the model's
    regular grid (device) and calls marching_cubes, which is the device isosurface; the returned Surface lies on the level
    set of the feature."""
    sys.path.insert(0, HERE)
    import refenv

    if refenv.activate() is None:
        pytest.skip("the unmodified reference is not available (baseline/_ref)")
    import LoopStructural as ls
    from LoopStructural.datasets import load_claudius

    import loopstructural_b200.plugin as b200

    b200.install()
    try:
        data, bb = load_claudius()
        model = ls.GeologicalModel(bb[0, :], bb[1, :])
        model.data = data
        model.create_and_add_foliation("strati", nelements=2e4, solver="b200")
        model.update()
        feat = model["strati"]
        vals = feat.evaluate_value(model.regular_grid(nsteps=(10, 10, 10)))
        level = float(np.nanmedian(vals))
        surfaces = feat.surfaces([level])
        assert len(surfaces) == 1
        s = surfaces[0]
        verts, tris = np.asarray(s.vertices), np.asarray(s.triangles)
        assert len(verts) > 100 and len(tris) > 100 and tris.max() < len(verts)
        f = feat.evaluate_value(verts)
        rng_ = float(np.nanmax(vals) - np.nanmin(vals))
        ok = np.isfinite(f)
        assert ok.mean() > 0.99
        # vertices interpolate the samples of the model's regular grid linearly along its edges; the field in between is
        # trilinear on the (finer) support: off by a fraction of the change across one lattice cell
        assert np.abs(f[ok] - level).max() < 0.05 * rng_ and np.median(np.abs(f[ok] - level)) < 0.005 * rng_
    finally:
        b200.uninstall()
