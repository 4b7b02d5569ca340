"""Per-CTA timeline of one launch of the regularisation operator (lsb200_debug_cta_trace): which role ends last, how long
the SMs sit idle.  usage: cta_trace.py fdi N [k0 k1] | p1 NCELLS"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loopstructural_b200 import Engine, engine as en  # noqa: E402

kind, n = sys.argv[1], int(sys.argv[2])
slab = tuple(int(v) for v in sys.argv[3:5]) if len(sys.argv) >= 5 else None
if kind == "fdi":
    e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1), slab=slab)
    e.set_regularisation({})
else:
    e = Engine([n + 1] * 3, np.zeros(3), np.ones(3) / n, kind="tetmesh")
    e.set_p1_regularisation(1.0)
pts = np.random.default_rng(0).uniform(0.01, 0.99, (1000, 3))
if slab:
    pts[:, 2] = (slab[0] + pts[:, 2] * (slab[1] - slab[0])) / (n - 1)
e.add_point_rows(en.ROWS_VALUE, np.column_stack([pts, pts[:, 2], np.ones(len(pts))]), 1.0)
e.finalize()
tr = e.cta_trace().astype(np.int64)
t0 = tr[:, 0].min()
start, end, sm = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2]
dur = end - start
print(f"{kind} {n} slab={slab}: {len(tr)} CTAs on {len(set(sm))} SMs, launch span {end.max() / 1e3:.1f} us")
print(f"  CTA duration us: min {dur.min() / 1e3:.1f} median {np.median(dur) / 1e3:.1f} max {dur.max() / 1e3:.1f}; "
      f"start us: median {np.median(start) / 1e3:.1f} max {start.max() / 1e3:.1f}")
# groups of consecutive block indices with similar duration (roles are contiguous in block-index order)
edges = [0]
for name, cnt in e.plan_roles() if hasattr(e, "plan_roles") else []:
    edges.append(edges[-1] + cnt)
step = max(1, len(tr) // 16)
for a in range(0, len(tr), step):
    b = min(len(tr), a + step)
    print(f"  CTAs {a:5d}-{b - 1:5d}: start {start[a:b].mean() / 1e3:6.1f} us, duration {dur[a:b].mean() / 1e3:6.1f} "
          f"(max {dur[a:b].max() / 1e3:6.1f}), end {end[a:b].mean() / 1e3:6.1f} (max {end[a:b].max() / 1e3:6.1f})")
per_sm_end = np.array([end[sm == s].max() for s in sorted(set(sm))])
per_sm_busy = np.array([len(np.flatnonzero(sm == s)) for s in sorted(set(sm))])
print(f"  per SM: CTAs min {per_sm_busy.min()} max {per_sm_busy.max()}; last end us: min {per_sm_end.min() / 1e3:.1f} "
      f"mean {per_sm_end.mean() / 1e3:.1f} max {per_sm_end.max() / 1e3:.1f}")
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"trace_{kind}_{n}.npy"), tr)
e.close()
