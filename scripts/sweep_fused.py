"""Sweep of the k_fdi_fused plan knobs on one GPU: box configuration (TYT), box chunk length, strip chunk length.
Each point runs in a process of its own (the knobs are read once per process).  usage: sweep_fused.py n [k0 k1]"""
import itertools
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
n = int(sys.argv[1])
slab = sys.argv[2:4]
code = f"""
import sys, json, os, numpy as np
sys.path.insert(0, {os.path.dirname(HERE)!r})
from loopstructural_b200 import Engine, engine as en
n = {n}
slab = {tuple(int(v) for v in slab) if slab else None}
e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1), slab=slab)
e.set_regularisation({{}})
pts = np.random.default_rng(0).uniform(0, 1, (1000, 3))
if slab: pts[:, 2] = (slab[0] + pts[:, 2] * (slab[1] - slab[0])) / (n - 1)
e.add_point_rows(en.ROWS_VALUE, np.column_stack([pts, pts[:, 2], np.ones(len(pts))]), 1.0)
e.finalize()
print(json.dumps({{"us": 1e6 * e.time_kernel(0, 20)}}))
"""
nzl = (int(slab[1]) - int(slab[0])) if slab else n
czs = sorted({max(8, nzl // k) for k in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12)})
sczs = [8, 12, 16, 24, 32, 48]
best = None
for cfg, cz, scz in itertools.product((4, 3), czs, sczs):
    env = dict(os.environ, LSB200_FUSED_CFG=str(cfg), LSB200_FUSED_CZ=str(cz), LSB200_FUSED_SCZ=str(scz))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    try:
        us = json.loads(r.stdout.strip().splitlines()[-1])["us"]
    except Exception:
        print("failed", cfg, cz, scz, r.stderr[-200:])
        continue
    print(f"n={n} slab={slab} cfg={cfg} cz={cz} scz={scz}: {us:.1f} us", flush=True)
    if best is None or us < best[0]:
        best = (us, cfg, cz, scz)
print("BEST", n, slab, best)
