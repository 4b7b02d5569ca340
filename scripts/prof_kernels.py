"""Short driver for ncu captures: builds the C2 problem and runs a few CG iterations / kernel reps.
usage: python scripts/prof_kernels.py [n] [which] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loopstructural_b200 import Engine, engine as en
from bench import synthetic_points

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1))
e.set_regularisation({})
value, gradient = synthetic_points(100_000, 10_000)
e.add_point_rows(en.ROWS_VALUE, value, 1.0)
e.add_point_rows(en.ROWS_GRADIENT, gradient, 1.0)
e.finalize()
t = e.time_kernel(which, reps)
print(f"n={n} which={which}: {t*1e6:.1f} us per launch")
