"""Where the end-to-end time of one public-API solve goes (host buffers in, host c out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import synthetic_points
from loopstructural_b200.interpolators import B200FiniteDifferenceInterpolator, Support

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
value, gradient = synthetic_points(100_000, 10_000)
it = B200FiniteDifferenceInterpolator(Support(np.zeros(3), [n - 1] * 3, np.ones(3) / (n - 1)))
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); it.set_value_constraints(value); it.set_gradient_constraints(gradient)
    t1 = time.perf_counter(); it.setup_interpolator()
    t2 = time.perf_counter(); it.solve_system("b200")
    t3 = time.perf_counter()
    t = it.get_last_solve_timing()
    print(f"rep {rep}: set {1e3*(t1-t0):.1f} ms, setup_interpolator {1e3*(t2-t1):.1f} ms, solve_system {1e3*(t3-t2):.1f} ms "
          f"(device solve {1e3*t['solve_seconds']:.1f}, finalize {1e3*t['preprocess_seconds']:.1f}, {t['iterations']} its)")
