"""One configuration of the FDI matvec, timed in a process of its own (launch-plan switches: LSB200_TMA, LSB200_FUSED_CFG,
LSB200_FUSED_CZ, LSB200_FUSED_SCZ, LSB200_STENCIL_MAIN).  Prints one JSON line:
matvec and whole-iteration time (CUDA events on the engine's stream, lsb200_time_kernel) and two checksums of
q = N p for a fixed p, so that the caller can compare the numerics of the configurations with each other.

    LSB200_TMA=0 python scripts/matvec_experiment.py 256
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loopstructural_b200 import Engine, engine as en  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1))
    e.set_regularisation({})
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1, (1000, 3))
    e.add_point_rows(en.ROWS_VALUE, np.column_stack([pts, pts[:, 2], np.ones(len(pts))]), 1.0)
    e.finalize()
    p = rng.standard_normal(e.local_dof)
    q = e.apply_normal(p, parts=en.PART_REGULARISATION | en.PART_RIDGE)
    probe = np.cos(np.arange(q.size) * 0.37)
    out = {"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("LSB200_")},
           "matvec_us": 1e6 * e.time_kernel(0, 20), "iteration_us": 1e6 * e.time_kernel(1, 20),
           "q_dot_probe": float(q @ probe), "q_norm": float(np.linalg.norm(q)), "p_dot_q": float(p @ q)}
    out["matvec_gbs"] = 16.0 * e.local_dof / out["matvec_us"] / 1e3
    out["plan"] = e.describe_plan()
    e.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
