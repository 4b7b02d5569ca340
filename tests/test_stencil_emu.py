"""The FDI regularisation operator -- k_fdi_fused (one launch: box tiles, z faces, x / y strips) and k_fdi_normal,
unmodified from loopstructural_b200/csrc -- run on the CPU by tests/hostsim (one OS thread per CUDA thread, see cuda_emu.h)
with the launch plan csrc/api.cu uses (csrc/fdi_plan.cuh), against the oracle's explicit operator.  What this adds to the
-m gpu tests: z SLABS.  A slab handle's halo planes are filled here as the NCCL exchange would fill them, so every
rank of a 2 / 3 / 4 / 8-way partition (ranks that own only the lower face, only the upper face, or no face) is
checked without a multi-GPU box."""
import ctypes as C
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

from loopstructural_b200 import partition
from loopstructural_b200.engine import effective_stencil_weights
from oracle import fdi_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "stencil_emu.cpp")
LIB = os.path.join(HERE, "hostsim", "libstencil_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "loopstructural_b200", "csrc")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emu():
    gxx = shutil.which("g++")
    if gxx is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("g++ / CUDA headers not available")
    deps = [SRC, os.path.join(HERE, "hostsim", "cuda_emu.h")] + glob.glob(os.path.join(CSRC, "*.cuh"))
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run([gxx, "-O1", "-std=c++20", "-pthread", "-mfma", "-w", "-DLSB_EMU", "-I/usr/local/cuda/include", "-fPIC",
                        # its own instantiations of the kernels, not the device stubs of liblsb200.so (RTLD_GLOBAL)
                        "-fvisibility=hidden", "-Wl,-Bsymbolic", "-shared", "-o", LIB, SRC], check=True, capture_output=True)
    return C.CDLL(LIB)


def run_normal(emu, nsteps, slab, weights, ridge2, p, cz=0):
    nx, ny, nz = (int(v) for v in nsteps)
    k0, k1 = slab
    dims = (C.c_int * 5)(nx, ny, nz, k0, k1)
    ew = np.ascontiguousarray(effective_stencil_weights(weights))
    q = np.empty(nx * ny * (k1 - k0))
    pq = C.c_double(0.0)
    info = (C.c_int * 9)()
    rc = emu.emu_fdi_normal(dims, ew.ctypes.data_as(dp), C.c_double(ridge2), cz, np.ascontiguousarray(p).ctypes.data_as(dp),
                            q.ctypes.data_as(dp), C.byref(pq), info)
    assert rc == 0, f"emulated launch failed ({rc}): -2 = a kernel wrote outside the owned planes"
    return q, pq.value, list(info)


ANISO = dict(dxx=1.3, dyy=0.7, dzz=2.1, dxy=0.9, dxz=1.7, dyz=0.4, dx=1.9, dy=0.6, dz=1.1)


@pytest.mark.parametrize("nsteps,world", [((9, 7, 13), 1), ((15, 15, 15), 3), ((12, 30, 5), 1), ((40, 36, 30), 2)])
def test_single_kernel_operator_on_every_slab_matches_the_oracle(emu, nsteps, world):
    """k_fdi_normal (grids too small to have a box, and LSB200_STENCIL_MAIN=-1)."""
    grid = fo.Grid(np.array(nsteps), np.zeros(3), np.ones(3))
    rng = np.random.default_rng(5)
    p = rng.standard_normal(grid.n_nodes)
    ref = fo.regularisation_normal_matvec(grid, ANISO, p) + 0.01 * p
    plane = nsteps[0] * nsteps[1]
    bounds = partition.slab_bounds(nsteps[2], world) if world > 1 else [(0, nsteps[2])]
    for k0, k1 in bounds:
        q, pq, info = run_normal(emu, nsteps, (k0, k1), ANISO, 0.01, p, cz=4)
        sl = slice(k0 * plane, k1 * plane)
        assert np.abs(q - ref[sl]).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)
        assert abs(pq - float(p[sl] @ ref[sl])) <= 1e-10 * abs(float(p @ ref))


# ---- the single-launch plan (k_fdi_fused): box tiles with y edges and z faces, x strips -------------------------------
def run_fused(emu, nsteps, slab, weights, ridge2, p, cfg=0, cz=0, scz=0, sm_count=148):
    nx, ny, nz = (int(v) for v in nsteps)
    k0, k1 = slab
    dims = (C.c_int * 5)(nx, ny, nz, k0, k1)
    ew = np.ascontiguousarray(effective_stencil_weights(weights))
    q = np.empty(nx * ny * (k1 - k0))
    pq = C.c_double(0.0)
    info = (C.c_int * 9)()
    rc = emu.emu_fdi_fused(dims, ew.ctypes.data_as(dp), C.c_double(ridge2), cfg, cz, scz, sm_count,
                           np.ascontiguousarray(p).ctypes.data_as(dp), q.ctypes.data_as(dp), C.byref(pq), info)
    assert rc == 0, f"emulated fused launch failed ({rc}): -2 = wrote outside the owned planes, -3 = no plan"
    return q, pq.value, list(info)


def check_fused(emu, nsteps, world, weights, cfg=0, cz=0, scz=0, sm_count=148, ridge2=0.09, seed=None):
    grid = fo.Grid(np.array(nsteps), np.zeros(3), np.ones(3))
    rng = np.random.default_rng(sum(nsteps) + world if seed is None else seed)
    p = rng.standard_normal(grid.n_nodes)
    ref = fo.regularisation_normal_matvec(grid, weights, p) + ridge2 * p
    plane = nsteps[0] * nsteps[1]
    bounds = partition.slab_bounds(nsteps[2], world) if world > 1 else [(0, nsteps[2])]
    scale = max(np.abs(ref).max(), 1.0)
    total_pq = 0.0
    for k0, k1 in bounds:
        q, pq, info = run_fused(emu, nsteps, (k0, k1), weights, ridge2, p, cfg=cfg, cz=cz, scz=scz, sm_count=sm_count)
        sl = slice(k0 * plane, k1 * plane)
        assert not np.isnan(q).any(), f"slab {k0, k1}: {int(np.isnan(q).sum())} owned nodes never written (plan {info})"
        err = np.abs(q - ref[sl])
        worst = np.unravel_index(int(err.argmax()), (k1 - k0, nsteps[1], nsteps[0]))
        assert err.max() <= 1e-12 * scale, f"slab {k0, k1} plan {info}: max error {err.max():.3e} at (k,j,i) = {worst}"
        assert abs(pq - float(p[sl] @ ref[sl])) <= 1e-10 * abs(float(p @ ref))
        total_pq += pq
    assert abs(total_pq - float(p @ ref)) <= 1e-10 * abs(float(p @ ref))


@pytest.mark.parametrize("nsteps,world,weights,cfg,cz,scz", [
    ((40, 36, 30), 1, {}, 4, 0, 0),
    ((40, 36, 30), 1, ANISO, 3, 0, 0),
    ((40, 36, 30), 2, ANISO, 4, 0, 0),      # rank 0 owns only the lower face, rank 1 only the upper one
    ((40, 36, 30), 3, ANISO, 3, 5, 7),      # a middle rank without a face, several chunks, strips chunked differently
    ((33, 18, 48), 8, ANISO, 4, 0, 0),      # odd pitch (8-byte staging), 6 planes per rank
    ((18, 41, 16), 8, {}, 3, 3, 4),         # 2 planes per rank
    ((70, 20, 9), 1, ANISO, 4, 0, 0),       # fewer than 10 planes: the two z-face tables overlap
    ((70, 20, 9), 2, ANISO, 3, 0, 0),
    ((134, 17, 12), 1, ANISO, 4, 4, 0),     # three box tiles in x, odd ny
    ((16, 16, 6), 1, ANISO, 0, 0, 0),       # the smallest grid that has a plan
])
def test_fused_operator_on_every_slab_matches_the_oracle(emu, nsteps, world, weights, cfg, cz, scz):
    check_fused(emu, nsteps, world, weights, cfg=cfg, cz=cz, scz=scz)


def test_fused_operator_random_grids_and_plans(emu):
    rng = np.random.default_rng(77)
    cases = [(nx, ny, nz) for nx in (16, 17, 19, 66, 71, 131) for ny in (16, 18, 19, 33, 50) for nz in (6, 7, 8, 11, 17, 23)]
    rng.shuffle(cases)
    for nx, ny, nz in cases[:12]:
        w = ANISO if rng.random() < 0.7 else {}
        world = min(int(rng.integers(1, max(2, nz // 2))), nz // 2)
        check_fused(emu, (nx, ny, nz), world, w, cfg=int(rng.choice([0, 3, 4])), cz=int(rng.choice([0, 3, 4, 7])),
                    scz=int(rng.choice([0, 4, 9])), sm_count=int(rng.choice([1, 4, 148])), seed=int(rng.integers(1 << 30)))
