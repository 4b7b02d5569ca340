"""Matvec / iteration time of ONE z-slab of a 1024 x 1024 x 1024 grid on one GPU (no neighbours: halos stay zero):
an end rank (owns a z face) against a middle rank.  usage: python scripts/prof_slab.py [n] [planes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loopstructural_b200 import Engine, engine as en
from bench import synthetic_points
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 128
for name, slab in (("end rank", (0, planes)), ("middle rank", (planes, 2 * planes))):
    e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1), slab=slab)
    e.set_regularisation({})
    v, g = synthetic_points(125_000, 12_500)
    v[:, 2] = (slab[0] + v[:, 2] * planes) / (n - 1); g[:, 2] = (slab[0] + g[:, 2] * planes) / (n - 1)
    e.add_point_rows(en.ROWS_VALUE, v, 1.0); e.add_point_rows(en.ROWS_GRADIENT, g, 1.0)
    e.finalize()
    dof = e.local_dof
    tm, ti = e.time_kernel(0, 10), e.time_kernel(1, 10)
    print(f"{n}x{n} x planes {slab} ({name}): matvec {tm*1e6:.1f} us = {16*dof/tm/1e9:.0f} GB/s, iteration {ti*1e6:.1f} us = {104*dof/ti/1e9:.0f} GB/s; {e.describe_plan()}")
    e.close()
