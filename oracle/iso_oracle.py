"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the lattice isosurface of csrc/iso.cuh (marching tetrahedra, six Kuhn
tetrahedra per cell, one vertex per crossed lattice edge): what the device entry point lsb200_isosurface must return, entry
for entry.  It stands where skimage.measure.marching_cubes stands in the reference (LoopStructural/utils/_surface.py:139-147);
scikit-image (a third-party dependency, absent from this image and from /root/reference) triangulates cubes with Lewiner's
tables, which cannot be consulted here: PARITY UNPINNED against it.  What pins this file: analytic surfaces
(tests/test_iso_oracle.py: vertices on the level set of the interpolant, closed manifold, area of a sphere)."""
from __future__ import annotations

import itertools

import numpy as np

PERMS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


def tet_corners(t):
    cur = [0, 0, 0]
    out = [tuple(cur)]
    for a in PERMS[t]:
        cur[a] = 1
        out.append(tuple(cur))
    return out


def marching_tetrahedra(volume, level, spacing=(1.0, 1.0, 1.0)):
    """(verts, faces, normals): see csrc/iso.cuh.  Pure-Python loop over the crossed tetrahedra: small volumes only."""
    v = np.asarray(volume, dtype=float)
    n0, n1, n2 = v.shape
    sp = np.asarray(spacing, dtype=float)
    inside = v >= level
    fin = np.isfinite(v)

    def node(i):
        return (i[0] * n1 + i[1]) * n2 + i[2]

    # crossed edges in edge-index order: edge = 7 * node + type, direction bits of type + 1
    edges = {}
    flags = []
    for t in range(7):
        d = ((t + 1) & 1, ((t + 1) >> 1) & 1, ((t + 1) >> 2) & 1)
        a = (slice(0, n0 - d[0]), slice(0, n1 - d[1]), slice(0, n2 - d[2]))
        b = (slice(d[0], n0), slice(d[1], n1), slice(d[2], n2))
        cross = fin[a] & fin[b] & (inside[a] != inside[b])
        idx = np.argwhere(cross)
        for i in idx:
            flags.append((7 * node(i) + t, tuple(i), d))
    flags.sort()
    verts = np.empty((len(flags), 3))
    normals = np.empty((len(flags), 3))

    def grad(i):
        g = np.zeros(3)
        nn = (n0, n1, n2)
        for a in range(3):
            lo, hi = list(i), list(i)
            if i[a] > 0:
                lo[a] -= 1
            if i[a] + 1 < nn[a]:
                hi[a] += 1
            vl, vh, vc = v[tuple(lo)], v[tuple(hi)], v[tuple(i)]
            if not np.isfinite(vl):
                vl, lo[a] = vc, i[a]
            if not np.isfinite(vh):
                vh, hi[a] = vc, i[a]
            dist = (hi[a] - lo[a]) * sp[a]
            g[a] = (vh - vl) / dist if dist > 0 else 0.0
        return g

    for k, (e, i, d) in enumerate(flags):
        edges[e] = k
        j = (i[0] + d[0], i[1] + d[1], i[2] + d[2])
        t = (level - v[i]) / (v[j] - v[i])
        verts[k] = (np.array(i) + t * np.array(d)) * sp
        ga, gb = grad(i), grad(j)
        nr = -(ga + t * (gb - ga))
        ln = np.linalg.norm(nr)
        normals[k] = nr / ln if ln > 0 else 0.0

    faces = []
    for i0, i1, i2 in itertools.product(range(n0 - 1), range(n1 - 1), range(n2 - 1)):
        cube = v[i0:i0 + 2, i1:i1 + 2, i2:i2 + 2]
        if np.all(inside[i0:i0 + 2, i1:i1 + 2, i2:i2 + 2]) or not np.any(inside[i0:i0 + 2, i1:i1 + 2, i2:i2 + 2] & np.isfinite(cube)):
            if np.all(np.isfinite(cube)):
                continue
        for t in range(6):
            c = tet_corners(t)
            val = [v[i0 + o[0], i1 + o[1], i2 + o[2]] for o in c]
            if not all(np.isfinite(val)):
                continue
            ins = [s for s in range(4) if val[s] >= level]
            outs = [s for s in range(4) if val[s] < level]
            if not ins or not outs:
                continue

            def ev(a, b):
                lo, hi = (a, b) if a < b else (b, a)
                dd = tuple(c[hi][k] - c[lo][k] for k in range(3))
                ty = (dd[0] | (dd[1] << 1) | (dd[2] << 2)) - 1
                return edges[7 * node((i0 + c[lo][0], i1 + c[lo][1], i2 + c[lo][2])) + ty]

            def cut(a, b):
                tt = (level - val[a]) / (val[b] - val[a])
                return (np.array(c[a]) + tt * (np.array(c[b]) - np.array(c[a]))) * sp

            direction = (np.mean([c[s] for s in outs], axis=0) - np.mean([c[s] for s in ins], axis=0)) * sp

            def oriented(tri, pairs):
                p = [cut(a, b) for a, b in pairs]
                if np.dot(np.cross(p[1] - p[0], p[2] - p[0]), direction) < 0.0:
                    tri = [tri[0], tri[2], tri[1]]
                return tri

            if len(ins) in (1, 3):
                lone = ins[0] if len(ins) == 1 else outs[0]
                others = outs if len(ins) == 1 else ins
                pairs = [(lone, o) for o in others]
                faces.append(oriented([ev(a, b) for a, b in pairs], pairs))
            else:
                a0, a1 = ins
                b0, b1 = outs
                for pairs in ([(a0, b0), (a0, b1), (a1, b1)], [(a0, b0), (a1, b1), (a1, b0)]):
                    faces.append(oriented([ev(a, b) for a, b in pairs], pairs))
    return verts, np.array(faces, dtype=np.int32).reshape(-1, 3), normals
