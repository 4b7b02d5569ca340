"""Seeded inputs of the full-size BASELINE configurations (SURVEY.md section 8 d), shared by the generator of the golden
fixtures (make_golden_fullsize.py, build container) and by the device tests, so that a fixture need not store them."""
import numpy as np


def field(p):
    return p[:, 2] + 0.1 * np.sin(2 * np.pi * p[:, 0]) + 0.05 * np.cos(2 * np.pi * p[:, 1])


def gradient_of_field(p):
    return np.stack([0.2 * np.pi * np.cos(2 * np.pi * p[:, 0]), -0.1 * np.pi * np.sin(2 * np.pi * p[:, 1]), np.ones(p.shape[0])], axis=1)


def c2_points(n, seed=0):
    """configs[1] scaled to an n^3-node grid: 1e5 value + 1e4 gradient points at 256^3, uniform in the unit cube."""
    nv, ng = max(100, int(1e5 * (n / 256.0) ** 3)), max(10, int(1e4 * (n / 256.0) ** 3))
    rng = np.random.default_rng(seed)
    pv = rng.uniform(0, 1, (nv, 3))
    pg = rng.uniform(0, 1, (ng, 3))
    return np.column_stack([pv, field(pv), np.ones(nv)]), np.column_stack([pg, gradient_of_field(pg), np.ones(ng)])


def c3_points(n_normal=100_000, n_tangent=100_000, n_value=100, seed=0):
    """configs[2]: 1e5 normal + 1e5 tangent (grad f x e_x) + 100 value points in [0.01, 0.99]^3."""
    rng = np.random.default_rng(seed)
    pn = rng.uniform(0.01, 0.99, (n_normal, 3))
    pt = rng.uniform(0.01, 0.99, (n_tangent, 3))
    pv = rng.uniform(0.01, 0.99, (n_value, 3))
    value = np.column_stack([pv, field(pv), np.ones(n_value)])
    normal = np.column_stack([pn, gradient_of_field(pn), np.ones(n_normal)])
    tangent = np.column_stack([pt, np.cross(gradient_of_field(pt), [1.0, 0.0, 0.0]), np.ones(n_tangent)])
    return value, normal, tangent


def probe_points(n, seed, lo=-0.02, hi=1.02):
    return np.random.default_rng(seed).uniform(lo, hi, (n, 3))
