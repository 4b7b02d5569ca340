"""Shared loaders for the golden fixtures (tests only)."""
import os

import numpy as np
from scipy import sparse

from oracle import fdi_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def weights_of(z, kk="weights_keys", kv="weights_vals"):
    return {str(k): float(v) for k, v in zip(z[kk], z[kv])}


def ref_family(z, name):
    shape = tuple(z[f"{name}/shape"])
    m = sparse.coo_matrix((z[f"{name}/vals"], (z[f"{name}/rows"], z[f"{name}/cols"])), shape=shape).tocsr()
    return m, z[f"{name}/b"], z[f"{name}/w"]


def solve_case(z):
    """grid, weights, oracle data families, rtol, atol of a fdi_solve_*.npz fixture."""
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    kw = weights_of(z, "kw_keys", "kw_vals")
    weights = dict(fo.DEFAULT_WEIGHTS)
    if "regularisation" in kw:
        for k in ("dxy", "dyz", "dxz", "dxx", "dyy", "dzz"):
            weights[k] = kw["regularisation"]
    weights.update({k: v for k, v in kw.items() if k != "regularisation"})
    fams = []
    if "data_normal" in z:
        fams += fo.normal_rows(grid, z["data_normal"], weights["npw"])
    if "data_gradient" in z:
        fams += fo.gradient_rows(grid, z["data_gradient"], weights["gpw"])
    fams.append(fo.value_rows(grid, z["data_value"], weights["cpw"]))
    if "data_tangent" in z:
        fams.append(fo.orthogonal_rows(grid, z["data_tangent"][:, :3], z["data_tangent"][:, 3:6], weights["tpw"]))
    rtol = 1e-5 if float(z["rtol"]) < 0 else float(z["rtol"])
    atol = 0.0 if float(z["tol"]) < 0 else float(z["tol"])
    return grid, weights, fams, rtol, atol


def engine_for(grid, weights=None, **opts):
    from loopstructural_b200 import Engine

    e = Engine(grid.nsteps, grid.origin, grid.step)
    e.set_regularisation(weights or {})
    if opts:
        e.set_options(**opts)
    return e


def add_point_data(e, z, weights, prefix="data_"):
    """Feed the fixture's constraint points through the device-side row generation (K3)."""
    from loopstructural_b200 import engine as en

    n = 0
    if f"{prefix}normal" in z:
        n += e.add_point_rows(en.ROWS_NORMAL, z[f"{prefix}normal"], weights["npw"])
    if f"{prefix}gradient" in z:
        n += e.add_point_rows(en.ROWS_GRADIENT, z[f"{prefix}gradient"], weights["gpw"])
    n += e.add_point_rows(en.ROWS_VALUE, z[f"{prefix}value"], weights["cpw"])
    if f"{prefix}tangent" in z:
        t = np.array(z[f"{prefix}tangent"], dtype=float)
        t[:, 6] = 0.0  # column 6 is the row target b for tangent rows
        n += e.add_point_rows(en.ROWS_TANGENT, t, weights["tpw"])
    return n


def fault_case(z):
    """fields, profile, displacement, steps, fault_offset of a fault_*.npz fixture (oracle objects)."""
    from oracle import fault_oracle as fa

    fields = [fa.FrameField(fo.Grid.from_cells(z[f"f{i}_ncells"], z[f"f{i}_origin"], z[f"f{i}_step"]), z[f"f{i}_c"],
                            z[f"f{i}_w2l"], "tetmesh" if f"f{i}_kind" in z and int(z[f"f{i}_kind"]) == 1 else "grid")
              for i in range(3)]

    def p1(row):
        return fa.Profile1D(int(row[0]), tuple(row[1:5]), (row[5], row[6]), float(row[7]), float(row[8]))

    axes = None
    if int(z["has_function"]):
        tab, comp = z["profile_table"], z["profile_composite"]
        axes = tuple(fa.AxisProfile(p1(tab[a, 0]), p1(tab[a, 1]) if comp[a] else None) for a in range(3))
    return fields, fa.FaultProfile(axes, float(z["profile_scale"])), float(z["displacement"]), int(z["steps"]), float(z["fault_offset"])
