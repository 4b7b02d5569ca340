"""Small-grid behaviour (BASELINE configs[0] size: Claudius, 55x68x47 nodes): iteration time vs launch overhead."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loopstructural_b200 import Engine, engine as en
from bench import synthetic_points
e = Engine([55, 68, 47], np.zeros(3), np.array([1 / 54, 1 / 67, 1 / 46]))
e.set_regularisation({})
v, g = synthetic_points(21023, 25)
e.add_point_rows(en.ROWS_VALUE, v, 1.0); e.add_point_rows(en.ROWS_NORMAL, g, 1.0)
c, res = e.solve_cg(rtol=1e-5, atol=5.4e-7)
c, res = e.solve_cg(rtol=1e-5, atol=5.4e-7)
print(f"55x68x47: {res.iterations} iterations, solve {res.solve_seconds*1e3:.1f} ms, {res.solve_seconds/res.iterations*1e6:.1f} us/iteration, "
      f"{res.kernel_launches/max(res.iterations,1):.1f} launches/iteration; matvec {e.time_kernel(0,50)*1e6:.1f} us, iteration {e.time_kernel(1,50)*1e6:.1f} us")
