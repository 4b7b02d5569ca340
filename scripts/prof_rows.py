"""Data-row kernels: atomic vs deterministic path, with ncu-friendly repetition."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loopstructural_b200 import Engine, engine as en
from bench import synthetic_points
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for det in (False, True):
    e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1))
    e.set_regularisation({})
    e.set_options(deterministic=det)
    value, gradient = synthetic_points(100_000, 10_000)
    e.add_point_rows(en.ROWS_VALUE, value, 1.0)
    e.add_point_rows(en.ROWS_GRADIENT, gradient, 1.0)
    e.finalize()
    print(f"n={n} deterministic={det}: rows {e.time_kernel(2, 20)*1e6:.1f} us, iteration {e.time_kernel(1, 20)*1e6:.1f} us")
    e.close()
