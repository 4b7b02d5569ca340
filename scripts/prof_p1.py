"""Times the P1 path at BASELINE configs[2] size (117^3 cells): operator launch, CG iteration, solve, evaluation.
usage: python scripts/prof_p1.py [ncells]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loopstructural_b200.interpolators import B200P1Interpolator, Support

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 117
rng = np.random.default_rng(0)
f = lambda p: p[:, 2] + 0.1 * np.sin(2 * np.pi * p[:, 0]) + 0.05 * np.cos(2 * np.pi * p[:, 1])
g = lambda p: np.stack([0.2 * np.pi * np.cos(2 * np.pi * p[:, 0]), -0.1 * np.pi * np.sin(2 * np.pi * p[:, 1]), np.ones(len(p))], axis=1)
pn, pt, pv = (rng.uniform(0.01, 0.99, (n, 3)) for n in (100_000, 100_000, 100))
it = B200P1Interpolator(Support(np.zeros(3), [nc] * 3, np.ones(3) / nc))
it.set_normal_constraints(np.column_stack([pn, g(pn), np.ones(len(pn))]))
it.set_tangent_constraints(np.column_stack([pt, np.cross(g(pt), [1.0, 0.0, 0.0]), np.ones(len(pt))]))
it.set_value_constraints(np.column_stack([pv, f(pv), np.ones(len(pv))]))
t0 = time.perf_counter(); it.setup_interpolator(cgw=0.1); t1 = time.perf_counter()
it.solve_system("b200"); t = it.get_last_solve_timing()
e = it.engine
print(f"P1 {nc}^3 cells ({5*nc**3} tets, {it.dof} nodes): setup {t1-t0:.3f} s, solve {t['solve_seconds']*1e3:.1f} ms, "
      f"{t['iterations']} CG iterations ({t['iterations']/t['solve_seconds']:.0f} it/s), info {t['info']}")
tm, ti = e.time_kernel(0, 20), e.time_kernel(1, 20)
print(f"  k_p1_normal {tm*1e6:.1f} us = {16*it.dof/tm/1e9:.0f} GB/s algorithmic (16 B/node); iteration {ti*1e6:.1f} us")
it.solve_system("b200")
q = rng.uniform(0, 1, (2_000_000, 3))
for name, fn in (("value", it.evaluate_value), ("gradient", it.evaluate_gradient)):
    fn(q[:1000]); t0 = time.perf_counter(); fn(q); dt = time.perf_counter() - t0
    print(f"  evaluate_{name}: {len(q)/dt/1e6:.1f} Mpt/s through host buffers")
