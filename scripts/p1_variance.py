import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import golden_helpers as H
from oracle import fdi_oracle as fo
from loopstructural_b200.interpolators import B200P1Interpolator, Support
for tag in ["unit12", "aniso", "unit16_tight"]:
    z = np.load(os.path.join(H.GOLDEN, f"p1_solve_{tag}.npz"))
    grid = fo.Grid.from_cells(z["ncells"], z["origin"], z["step"])
    out = []
    for rep in range(6):
        it = B200P1Interpolator(Support(grid.origin, grid.nsteps - 1, grid.step))
        it.apply_scaling_matrix = bool(z["scaling"])
        for key in ("value", "normal", "tangent"):
            if f"data_{key}" in z:
                getattr(it, f"set_{key}_constraints")(z[f"data_{key}"])
        it.setup_interpolator(**H.weights_of(z, "kw_keys", "kw_vals"))
        rtol = 1e-5 if float(z["rtol"]) < 0 else float(z["rtol"]); atol = 0.0 if float(z["tol"]) < 0 else float(z["tol"])
        it.solve_system("b200", solver_kwargs={"rtol": rtol, "atol": atol})
        t = it.get_last_solve_timing(); rng_ = float(z["c"].max() - z["c"].min())
        out.append((t["iterations"], float(np.linalg.norm(it.c - z["c"]) / np.sqrt(it.c.size) / rng_)))
        it.close()
    print(tag, "ref its", int(z["iterations"]), [(a, f"{b:.2e}") for a, b in out])
