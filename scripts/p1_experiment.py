"""P1 / TetMesh face-jump operator at BASELINE configs[2] size (117^3 cells = 8,008,065 tets), timed in a process of its
own (LSB200_P1_TILES / LSB200_P1_TILE_CFG are read once per process).  Prints one JSON line like scripts/matvec_experiment.py.

    LSB200_P1_TILE_CFG=1 python scripts/p1_experiment.py 117
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loopstructural_b200 import Engine, engine as en  # noqa: E402


def main():
    nc = int(sys.argv[1]) if len(sys.argv) > 1 else 117
    e = Engine([nc + 1] * 3, np.zeros(3), np.ones(3) / nc, kind="tetmesh")
    e.set_p1_regularisation(1.0)
    rng = np.random.default_rng(0)
    pts = rng.uniform(0.01, 0.99, (1000, 3))
    e.add_point_rows(en.ROWS_VALUE, np.column_stack([pts, pts[:, 2], np.ones(len(pts))]), 1.0)
    e.finalize()
    p = rng.standard_normal(e.local_dof)
    q = e.apply_normal(p, parts=en.PART_REGULARISATION | en.PART_RIDGE)
    probe = np.cos(np.arange(q.size) * 0.37)
    out = {"cells": nc, "env": {k: v for k, v in os.environ.items() if k.startswith("LSB200_")},
           "matvec_us": 1e6 * e.time_kernel(0, 20), "iteration_us": 1e6 * e.time_kernel(1, 20),
           "plan": e.describe_plan(), "q_dot_probe": float(q @ probe), "q_norm": float(np.linalg.norm(q)), "p_dot_q": float(p @ q)}
    out["matvec_gbs"] = 16.0 * e.local_dof / out["matvec_us"] / 1e3
    e.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
