#!/bin/bash
# Round 2, GPU session AG (one B200): device isosurface (marching tetrahedra) against the oracle, through the plug-in
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_iso_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/ag_pytest_iso.txt 2>&1; tail -15 gpurun_out/ag_pytest_iso.txt
python - <<'PY'
import time, numpy as np
from loopstructural_b200.isosurface import marching_tetrahedra
for n in (128, 256, 400):
    ax = np.linspace(-1, 1, n); x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    vol = 0.7 - np.sqrt(x*x + y*y + z*z) + 0.1*np.sin(5*x)*np.cos(4*y)
    marching_tetrahedra(vol, 0.0)
    t0 = time.perf_counter(); v, f, nr, _ = marching_tetrahedra(vol, 0.0); dt = time.perf_counter() - t0
    print(f"isosurface {n}^3: {len(v)} vertices, {len(f)} triangles, {dt*1e3:.1f} ms host arrays in and out ({vol.size/dt/1e6:.0f} M samples/s)")
PY
