"""CG iteration time of small systems (BASELINE configs[0]-sized FDI grid, a 64^3 grid, configs[2]: P1 on 117^3 cells) and a
full solve of each, in a process of its own: LSB200_PERSIST=0 keeps the launch-per-phase loop, the default runs the
iterations of a batch inside one persistent launch (csrc/cg_small.cuh).  Prints one JSON line per system."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import fullsize_inputs as fi  # noqa: E402
from loopstructural_b200 import Engine, engine as en  # noqa: E402


def run(name, e, add):
    add(e)
    e.finalize()
    c, res = e.solve_cg(rtol=1e-5)
    t0 = time.perf_counter()
    c, res = e.solve_cg(rtol=1e-5)
    wall = time.perf_counter() - t0
    out = {"system": name, "env": {k: v for k, v in os.environ.items() if k.startswith("LSB200_")}, "dof": int(e.local_dof),
           "iterations": int(res.iterations), "info": int(res.info), "solve_ms": 1e3 * res.solve_seconds, "wall_ms": 1e3 * wall,
           "us_per_iteration_solve": 1e6 * res.solve_seconds / max(int(res.iterations), 1),
           "iteration_us": 1e6 * e.time_kernel(1, 64), "matvec_us": 1e6 * e.time_kernel(0, 20), "launches": int(res.kernel_launches),
           "checksum": float(np.dot(c, np.cos(np.arange(c.size) * 0.37))), "plan": e.describe_plan()}
    try:
        out["phases_us"] = {k: round(v, 2) for k, v in e.small_phases(64).items()}
    except Exception:
        pass  # launch-by-launch iteration
    print(json.dumps(out), flush=True)
    e.close()


def fdi_points(e, nv, ng):
    v, g = fi.c2_points(256)
    e.add_point_rows(en.ROWS_VALUE, v[:nv], 1.0)
    e.add_point_rows(en.ROWS_GRADIENT, g[:ng], 1.0)


def main():
    which = sys.argv[1:] or ["claudius", "fdi64", "p1"]
    if "claudius" in which:
        ns = np.array([55, 68, 47])
        e = Engine(ns, np.zeros(3), 1.0 / (ns - 1))
        e.set_regularisation({})
        run("fdi 55x68x47 (configs[0] size), 21000 rows", e, lambda e: fdi_points(e, 15000, 3000))
    if "fdi64" in which:
        e = Engine([64, 64, 64], np.zeros(3), np.ones(3) / 63)
        e.set_regularisation({})
        run("fdi 64^3, 2000 rows", e, lambda e: fdi_points(e, 1500, 250))
    if "p1" in which:
        nc = 117
        e = Engine([nc + 1] * 3, np.zeros(3), np.ones(3) / nc, kind="tetmesh")
        e.set_p1_regularisation(0.1)
        value, normal, tangent = fi.c3_points()

        def add(e):
            e.add_point_rows(en.ROWS_VALUE, value, 1.0)
            e.add_point_rows(en.ROWS_NORMAL, normal, 1.0)
            e.add_point_rows(en.ROWS_TANGENT, tangent, 1.0)

        run("p1 117^3 cells (configs[2])", e, add)


if __name__ == "__main__":
    main()
