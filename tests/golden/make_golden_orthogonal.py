"""Golden fixtures of the frame-orthogonality rows (SURVEY.md section 8 f, rank 2) from the UNMODIFIED reference:
the '<feature>_orthogonal' families that GeologicalFeatureBuilder.install_gradient_constraint puts into the second and
third coordinate of a fault frame (LoopStructural.datasets.load_intrusion through GeologicalModel), once with the
dataset's own origin (world -> local is a translation) and once shifted to a zero origin; each once with FDI frames
(StructuredGrid) and once with PLI frames (TetMesh: orth_p1_*.npz).

    python tests/golden/make_golden_orthogonal.py      (build container only)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401  (sets up the reference import path and stubs)

from LoopStructural import GeologicalModel  # noqa: E402
from LoopStructural.datasets import load_intrusion  # noqa: E402


def main():
    data, bb = load_intrusion()
    for tag, shift, itype in (("world", False, "FDI"), ("zero_origin", True, "FDI"), ("p1_world", False, "PLI"),
                              ("p1_zero_origin", True, "PLI")):
        d, b2 = data.copy(), bb.copy()
        if shift:
            d[["X", "Y", "Z"]] -= bb[0]
            b2 = bb - bb[0]
        model = GeologicalModel(b2[0, :], b2[1, :])
        model.data = d
        model.create_and_add_fault("fault", 500, nelements=2000, interpolatortype=itype, buffer=0.3)
        model.update()
        fault = model["fault"]
        out = {}
        for i in range(3):
            it = fault[i].interpolator
            sup = it.support
            out[f"f{i}_ncells"] = np.asarray(sup.nsteps_cells)
            out[f"f{i}_origin"] = np.asarray(sup.origin, float)
            out[f"f{i}_step"] = np.asarray(sup.step_vector, float)
            out[f"f{i}_c"] = np.array(it.c)
            out[f"f{i}_w2l"] = np.array(it.bounding_box.world_to_local_matrix, float)
        fams = []
        for i in (1, 2):
            names = {f.name: j for j, f in enumerate(fault.features[:3])}
            for other, (feat, w, _region, step, B) in fault[i].builder._orthogonal_features.items():
                key = f"t{i}_o{names[feat.name]}"
                fams.append(key)
                out.update({f"{key}/step": int(step), f"{key}/weight": float(w), f"{key}/B": float(B)})
                c = fault[i].interpolator.constraints.get(f"{feat.name}_orthogonal")
                if c is None:
                    # a NaN in the family (a barycentre outside the other feature's tet mesh): add_constraints_to_least_squares
                    # drops the whole family (LI/_discrete_interpolator.py:357-359)
                    out[f"{key}/dropped"] = np.array(1)
                    print(tag, key, "DROPPED by the reference")
                    continue
                m = c["matrix"].tocoo()
                out.update({f"{key}/rows": m.row, f"{key}/cols": m.col, f"{key}/vals": m.data, f"{key}/shape": np.array(m.shape),
                            f"{key}/b": np.asarray(c["b"], float), f"{key}/w": np.asarray(c["w"], float)})
                zero = int(np.sum(np.asarray(abs(c["matrix"]).sum(axis=1)).ravel() == 0))
                print(tag, key, "rows", m.shape[0], "zero rows", zero, "step", step, "w", w, "B", B)
        out["kind"] = np.array(itype)
        np.savez_compressed(os.path.join(HERE, f"orth_{tag}.npz"), families=np.array(fams), **out)


if __name__ == "__main__":
    main()
