"""Host-side setup of the P1 regularisation stencil (constant-size work, done once per support).

P1Interpolator.minimise_edge_jumps (LI/_p1interpolator.py:115-166) adds one row n . (grad phi_T1 - grad phi_T2) = 0
per interior shared face of the TetMesh, row-normalised over its 8 listed entries
(LI/_discrete_interpolator.py:361-366).  On the structured 5-tets-per-cell lattice
(LC/supports/_3d_structured_tetra.py:28-34) the row of R^T R that belongs to a node is determined by the node's
parity and by how many cells (0, 1, >= 2) lie on either side of it on each axis, and it reaches +-2 nodes.  So the
operator of a mesh of ANY size is a lookup into R^T R of a mesh of at most 6 cells per axis with the same step
vector and the same cell-count parity.  This module assembles that small R^T R (numpy, a few thousand faces) and
cuts it into the table layout of lsb200_set_p1_tables; csrc/p1.cuh (k_p1_normal, p1_rep) does the lookup.
"""
from __future__ import annotations

import numpy as np

_TET_EVEN = np.array([[7, 1, 2, 4], [6, 2, 4, 7], [5, 1, 4, 7], [0, 1, 2, 4], [3, 1, 2, 7]])
_TET_ODD = np.array([[0, 6, 5, 3], [7, 3, 5, 6], [4, 0, 5, 6], [2, 0, 3, 6], [1, 0, 3, 5]])
_CORNER = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)])
_I34 = np.array([[-1.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 1.0]])


def small_mesh_cells(ncells) -> np.ndarray:
    """Cells per axis of the representative mesh: the mesh itself up to 6 cells, else 5 or 6 with equal parity."""
    nc = np.asarray(ncells, dtype=np.int64).reshape(3)
    return np.where(nc <= 6, nc, np.where(nc % 2 == 1, 5, 6)).astype(np.int64)


def _small_gram(m: np.ndarray, step: np.ndarray) -> np.ndarray:
    """Dense R^T R (unit weight) of the m-cell mesh; node id = i + (m0+1) (j + (m1+1) k)."""
    n = m + 1
    zz, yy, xx = np.meshgrid(np.arange(m[2]), np.arange(m[1]), np.arange(m[0]), indexing="ij")
    cells = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    corner = cells[:, None, :] + _CORNER[None, :, :]
    gi = corner[..., 0] + n[0] * (corner[..., 1] + n[1] * corner[..., 2])
    even = cells.sum(axis=1) % 2 == 0
    tets = np.where(even[:, None, None], gi[:, _TET_EVEN], gi[:, _TET_ODD]).reshape(-1, 4)

    def xyz(ids):
        k = ids // (n[0] * n[1])
        j = (ids // n[0]) % n[1]
        i = ids % n[0]
        return np.stack([i, j, k], axis=-1) * step

    # interior faces: node triples that belong to two tets
    drop = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    faces = np.sort(tets[:, drop].reshape(-1, 3), axis=1)
    owner = np.repeat(np.arange(tets.shape[0]), 4)
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    fs, ow = faces[order], owner[order]
    same = np.all(fs[1:] == fs[:-1], axis=1)
    t1, t2, shared = ow[:-1][same], ow[1:][same], fs[:-1][same]
    sx = xyz(shared)
    normal = np.cross(sx[:, 1] - sx[:, 0], sx[:, 2] - sx[:, 0])

    def grads(t):
        v = xyz(tets[t])
        return np.linalg.inv(v[:, 1:, :] - v[:, :1, :]) @ _I34

    rows = np.hstack([np.einsum("ij,ijk->ik", normal, grads(t1)), -np.einsum("ij,ijk->ik", normal, grads(t2))])
    rows /= np.linalg.norm(rows, axis=1)[:, None]
    cols = np.hstack([tets[t1], tets[t2]])
    nn = int(np.prod(n))
    R = np.zeros((rows.shape[0], nn))
    np.add.at(R, (np.repeat(np.arange(rows.shape[0]), 8), cols.ravel()), rows.ravel())
    return R.T @ R


def regularisation_table(ncells, step):
    """(m, table) for lsb200_set_p1_tables: table[(k*(m1+1) + j)*(m0+1) + i][tap] = (R^T R)[node, node + offset]."""
    m = small_mesh_cells(ncells)
    step = np.asarray(step, dtype=np.float64).reshape(3)
    n = m + 1
    N = _small_gram(m, step)
    nn = int(np.prod(n))
    kk, jj, ii = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    table = np.zeros((nn, 125))
    covered = np.zeros_like(N, dtype=bool)
    node = np.arange(nn)
    t = 0
    for dz in range(-2, 3):
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                ok = (ii + dx >= 0) & (ii + dx < n[0]) & (jj + dy >= 0) & (jj + dy < n[1]) & (kk + dz >= 0) & (kk + dz < n[2])
                nb = (ii + dx) + n[0] * ((jj + dy) + n[1] * (kk + dz))
                table[node[ok], t] = N[node[ok], nb[ok]]
                covered[node[ok], nb[ok]] = True
                t += 1
    if np.any(N[~covered] != 0.0):
        raise AssertionError("P1 regularisation couples nodes more than two apart: table layout too small")
    return m, np.ascontiguousarray(table)
