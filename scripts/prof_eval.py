"""Evaluation throughput (K7) with device-resident points: random cloud and grid-ordered points, value and gradient.
usage: python scripts/prof_eval.py [n_nodes_per_axis] [n_points]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from loopstructural_b200 import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
npts = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
e = Engine([n, n, n], np.zeros(3), np.ones(3) / (n - 1))
e.set_regularisation({})
rng = np.random.default_rng(0)
e.set_c(rng.standard_normal(n ** 3))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
cloud = torch.rand((npts, 3), dtype=torch.float64, device=dev, generator=g)
m = round(npts ** (1 / 3))
ax = torch.linspace(0, 1, m, dtype=torch.float64, device=dev)
grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3).flip(-1).contiguous()  # x fastest
out1 = torch.empty(npts, dtype=torch.float64, device=dev)
out3 = torch.empty((npts, 3), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for name, pts in (("random cloud", cloud), ("grid ordered", grid)):
    k = pts.shape[0]
    for what, fn, out, bytes_pt in (("value", e.evaluate_value_device, out1, 96), ("gradient", e.evaluate_gradient_device, out3, 112)):
        for _ in range(2):
            fn(pts.data_ptr(), out.data_ptr(), k, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn(pts.data_ptr(), out.data_ptr(), k, st)
        b.record(); torch.cuda.synchronize()
        dt = a.elapsed_time(b) * 1e-3 / 5
        floor = 32 if what == "value" else 48
        print(f"{n}^3 nodes, {k:.3g} pts {name:13s} evaluate_{what:8s}: {dt*1e3:7.2f} ms = {k/dt/1e9:6.2f} Gpt/s, "
              f"{bytes_pt*k/dt/1e9:6.0f} GB/s at {bytes_pt} B/pt (gather counted), {floor*k/dt/1e9:5.0f} GB/s at the {floor} B/pt streaming floor")
