"""q = N p of one grid under the launch-plan switches of the environment against k_fdi_normal (LSB200_STENCIL_MAIN=-1) in
a child process: max |difference| / max |q|.  usage: plan_check.py nx ny nz"""
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def compute(ns):
    from loopstructural_b200 import Engine
    e = Engine(ns, np.zeros(3), 1.0 / (np.array(ns) - 1))
    e.set_regularisation(dict(dxx=1.3, dyy=0.7, dzz=2.1, dxy=0.9, dxz=1.7, dyz=0.4, dx=1.9, dy=0.6, dz=1.1))
    e.set_options(ridge_factor=0.3)
    p = np.random.default_rng(1).standard_normal(e.local_dof)
    q = e.apply_normal(p, parts=1 | 4)
    q2 = e.apply_normal(p, parts=1 | 4)
    plan = e.describe_plan()
    e.close()
    return q, q2, plan


if __name__ == "__main__":
    ns = [int(v) for v in sys.argv[1:4]]
    if os.environ.get("PLAN_CHECK_CHILD"):
        q, _, _ = compute(ns)
        np.save(os.environ["PLAN_CHECK_CHILD"], q)
        sys.exit(0)
    ref_path = f"/tmp/plan_check_ref_{os.getpid()}.npy"
    env = dict(os.environ, PLAN_CHECK_CHILD=ref_path, LSB200_STENCIL_MAIN="-1")
    subprocess.run([sys.executable, __file__] + [str(v) for v in ns], env=env, check=True)
    ref = np.load(ref_path)
    q, q2, plan = compute(ns)
    print(json.dumps({"grid": ns, "env": {k: v for k, v in os.environ.items() if k.startswith("LSB200_")},
                      "max_rel_diff": float(np.abs(q - ref).max() / np.abs(ref).max()),
                      "repeat_diff": float(np.abs(q - q2).max() / np.abs(ref).max()), "plan": plan}))
