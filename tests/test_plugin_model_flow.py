"""The plug-in inside the reference's real call chain on CPU: GeologicalModel.create_and_add_foliation(solver='b200')
-> GeologicalFeatureBuilder -> interpolator_map -> B200FiniteDifferenceInterpolator.setup_interpolator / solve_system /
evaluate.  The device engine is replaced BY THIS TEST with a stand-in that solves the same system with the oracle, so
what is checked is the plumbing: which families reach the engine, the weight conventions, ridge / scaling / tolerance
options, and that the model's answers equal the stock solver's.  (The engine itself is checked on the GPU.)"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

pytestmark = pytest.mark.skipif(not refenv.available(), reason="reference neither installed (baseline/_ref) nor mounted")


class OracleEngine:
    """Same surface as loopstructural_b200.engine.Engine, numerics from oracle/ (test stand-in only)."""

    instances = []

    def __init__(self, nsteps, origin, step, kind="grid", device=0, slab=None):
        from oracle import fdi_oracle as fo

        self.fo = fo
        self.kind = kind
        self.grid = fo.Grid(np.asarray(nsteps), np.asarray(origin, float), np.asarray(step, float))
        self.rows, self.weights, self.opts, self.c = [], None, {}, None
        OracleEngine.instances.append(self)

    def set_regularisation(self, weights):
        self.weights = dict(weights)

    def set_p1_regularisation(self, w):
        self.cgw = float(w)

    def set_options(self, **kw):
        self.opts = kw

    def clear_rows(self):
        self.rows = []
        self.device_families = 0
        self.rhs_override = None

    def set_rhs(self, rhs):
        self.rhs_override = np.array(rhs, dtype=float)

    def set_c(self, c):
        self.c = np.array(c, dtype=float)

    def add_rows(self, A, idx, b, w, normalise=True):
        assert normalise is False  # the reference has normalised them already
        self.rows.append(self.fo.Family(f"rows{len(self.rows)}", np.array(A, float), np.array(idx), np.array(b, float), np.array(w, float)))

    def add_point_rows(self, kind, data, weight):
        from oracle import p1_oracle as po

        assert self.kind == "tetmesh" and kind == 4  # LSB200_ROWS_ELEMENT_VECTOR: the only point rows the plug-in generates
        data = np.asarray(data, float)
        if data.shape[0] == 0:
            return 0
        b = data[:, 6]
        assert np.all(b == b[0])
        fam = po.element_vector_rows(self.grid, data[:, :3], data[:, 3:6], float(weight), float(b[0]), name=f"device_element{len(self.rows)}")
        self.device_families = getattr(self, "device_families", 0) + 1
        self.rows.append(fam)
        return fam.A.shape[0]

    def add_directional_rows(self, vectors, weight):
        assert self.kind == "grid"
        fams = self.fo.directional_families(self.grid, vectors, float(weight), name=f"device_directional{len(self.rows)}")
        self.device_families = getattr(self, "device_families", 0) + 1
        self.rows.extend(fams)
        return sum(f.A.shape[0] for f in fams)

    def add_face_vector_rows(self, nodes_a, nodes_b, vectors, weight):
        from oracle import p1_oracle as po

        assert self.kind == "tetmesh"
        fam = po.face_vector_rows(self.grid, np.asarray(nodes_a), np.asarray(nodes_b), vectors, float(weight), name=f"device_face{len(self.rows)}")
        self.device_families = getattr(self, "device_families", 0) + 1
        self.rows.append(fam)
        return fam.A.shape[0]

    def add_field_orthogonal_rows(self, other, step=10, weight=1.0, b=0.0, other_world_to_local=None):
        from oracle import fault_oracle as fa

        src = fa.FrameField(other.grid, other.c, other_world_to_local, other.kind)
        make = fa.p1_field_orthogonal_rows if self.kind == "tetmesh" else fa.field_orthogonal_rows
        fam = make(self.grid, src, step, weight, b, name=f"device_orthogonal{len(self.rows)}")
        self.device_families = getattr(self, "device_families", 0) + 1
        if fam is None:  # NaN gradient outside a tet-mesh source: the whole family is dropped, as in the reference
            return 0
        self.rows.append(fam)
        return fam.A.shape[0]

    def solve_cg(self, rtol=1e-5, atol=0.0, maxiter=None, x0=None, want_c=True, out=None):
        if self.kind == "tetmesh":
            from oracle import p1_oracle as po

            fams = self.rows + ([po.face_rows(self.grid, self.cgw)] if self.cgw else [])
        else:
            w = dict(self.fo.DEFAULT_WEIGHTS)
            w.update(self.weights)
            fams = self.rows + self.fo.regularisation_families(self.grid, w)
        self.solve_args = dict(rtol=rtol, atol=atol)
        if getattr(self, "rhs_override", None) is not None:
            # lsb200_set_rhs: the operator of the rows, the right-hand side of the caller, CG from x0 in the scaled variables
            A, b = self.fo.assemble(fams, self.grid.n_nodes)
            A, b, S = self.fo.preprocess(A, b, ridge_factor=self.opts.get("ridge_factor", 1e-8),
                                         apply_scaling=self.opts.get("apply_scaling", True))
            S = np.ones(self.grid.n_nodes) if S is None else S
            N = (A.T @ A).tocsr()
            res = self.fo.cg(lambda v: N @ v, S * self.rhs_override, rtol=rtol, atol=atol, maxiter=maxiter,
                             x0=None if x0 is None else np.asarray(x0) / S)
            self.c = S * res.x
            return self.c, SimpleNamespace(info=res.info, iterations=res.iterations, setup_seconds=0.0, solve_seconds=0.0,
                                           residual_norm=res.residual_norm, threshold=res.threshold, kernel_launches=0)
        c, res, _ = self.fo.solve_reference_path(fams, self.grid.n_nodes, rtol=rtol, atol=atol,
                                                 ridge_factor=self.opts.get("ridge_factor", 1e-8),
                                                 apply_scaling=self.opts.get("apply_scaling", True))
        self.c = c
        return c, SimpleNamespace(info=res.info, iterations=res.iterations, setup_seconds=0.0, solve_seconds=0.0,
                                  residual_norm=res.residual_norm, threshold=res.threshold, kernel_launches=0)

    def evaluate_value(self, xyz):
        if self.kind == "tetmesh":
            from oracle import p1_oracle as po

            xyz = np.asarray(xyz, float)
            out = np.zeros(len(xyz))  # rows with NaN coordinates give 0.0 (LI/_discrete_interpolator.py:1357-1361)
            good = ~np.any(np.isnan(xyz), axis=1)
            out[good] = po.evaluate_value(self.grid, xyz[good], self.c)
            return out
        return self.fo.evaluate_value(self.grid, np.asarray(xyz, float), self.c)

    def evaluate_gradient(self, xyz):
        if self.kind == "tetmesh":
            from oracle import p1_oracle as po

            return po.evaluate_gradient(self.grid, np.asarray(xyz, float), self.c)
        return self.fo.evaluate_gradient(self.grid, np.asarray(xyz, float), self.c)

    def evaluate_lattice(self, start, stop, n, order="F", gradient=False, world_to_local=None, out_ptr=None, stream=0):
        from oracle import fault_oracle as fa

        self.lattice_calls = getattr(self, "lattice_calls", 0) + 1
        return fa.evaluate_lattice(fa.FrameField(self.grid, self.c, world_to_local, self.kind), start, stop, n, order, gradient)

    def close(self):
        pass


@pytest.fixture()
def ls(monkeypatch):
    refenv.activate()
    import LoopStructural
    from loopstructural_b200 import engine, plugin

    from loopstructural_b200 import fault as fl

    monkeypatch.setattr(engine, "Engine", OracleEngine)
    monkeypatch.setattr(fl, "faulted_evaluate", lambda *a, **k: oracle_faulted_evaluate(*a, **k))
    OracleEngine.instances.clear()
    plugin.install()
    yield LoopStructural
    plugin.uninstall()


def _model(ls, **kw):
    from LoopStructural.datasets import load_claudius

    data, bb = load_claudius()
    model = ls.GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    model.create_and_add_foliation("strati", nelements=3e3, **kw)
    model.update()
    return model, bb


def test_foliation_through_the_model_with_the_b200_solver_matches_the_stock_solver(ls):
    model, bb = _model(ls, solver="b200")
    it = model["strati"].interpolator
    assert type(it).__name__ == "B200FiniteDifferenceInterpolator"
    assert it.latest_solve_timing["backend"] == "b200" and it.up_to_date
    eng = OracleEngine.instances[-1]
    # only data families reached the engine as rows; the stencil families arrived as nine effective weights
    assert sorted(it.constraints) == ["norm_0", "norm_1", "norm_2", "value"] and len(eng.rows) == 4
    assert eng.weights == {"dxx": 1.0, "dyy": 1.0, "dzz": 1.0, "dxy": 1.0, "dxz": 1.0, "dyz": 1.0, "dx": 1.0, "dy": 1.0, "dz": 1.0}
    assert eng.opts["ridge_factor"] == it.ridge_factor and eng.opts["apply_scaling"] is True
    # the model passes tol (-> atol) and no rtol: rtol stays scipy's 1e-5 (LI/_solver_strategy.py:59-64)
    assert eng.solve_args["rtol"] == 1e-5 and eng.solve_args["atol"] > 0
    rng = np.random.default_rng(1)
    pts = bb[0] + rng.uniform(0, 1, (50, 3)) * (bb[1] - bb[0])
    val, grad = model.evaluate_feature_value("strati", pts), model.evaluate_feature_gradient("strati", pts)
    c_b200 = np.array(it.c)
    field = model["strati"].scalar_field()  # regular-grid export: node values from the lattice entry point
    assert eng.lattice_calls == 1
    gnorm = model["strati"].gradient_norm_scalar_field()  # |grad f| on the same lattice: the gradient variant of the entry point
    assert eng.lattice_calls == 2

    from loopstructural_b200 import plugin

    plugin.uninstall()
    stock, _ = _model(ls)  # stock class, stock 'cg'
    its = stock["strati"].interpolator
    assert type(its).__name__ == "FiniteDifferenceInterpolator" and its.latest_solve_timing["backend"] == "python"
    rng_ = float(its.c.max() - its.c.min())
    assert np.linalg.norm(c_b200 - its.c) / np.sqrt(its.c.size) / rng_ <= 1e-9
    np.testing.assert_allclose(val, stock.evaluate_feature_value("strati", pts), rtol=0, atol=1e-9 * rng_)
    np.testing.assert_allclose(grad, stock.evaluate_feature_gradient("strati", pts), rtol=0, atol=1e-12)
    stock_field = stock["strati"].scalar_field()
    assert type(stock["strati"]).scalar_field.__qualname__ == "BaseFeature.scalar_field"
    for got, ref in ((field.properties, stock_field.properties), (field.cell_properties, stock_field.cell_properties)):
        np.testing.assert_allclose(np.asarray(got["strati"], float), np.asarray(ref["strati"], float), rtol=0, atol=1e-9 * rng_,
                                   equal_nan=True)
    stock_gnorm = stock["strati"].gradient_norm_scalar_field()
    assert type(stock["strati"]).gradient_norm_scalar_field.__qualname__ == "BaseFeature.gradient_norm_scalar_field"
    for got, ref in ((gnorm.properties, stock_gnorm.properties), (gnorm.cell_properties, stock_gnorm.cell_properties)):
        np.testing.assert_allclose(np.asarray(got["strati"], float), np.asarray(ref["strati"], float), rtol=0, atol=1e-11,
                                   equal_nan=True)


def test_model_without_the_b200_solver_is_untouched_by_the_plugin(ls):
    model, _ = _model(ls)  # plug-in installed, default solver
    it = model["strati"].interpolator
    assert type(it).__name__ == "B200FiniteDifferenceInterpolator"
    assert it.latest_solve_timing["backend"] == "python" and "dxx" in it.constraints and not OracleEngine.instances


class OracleFaultFrame:
    """Same surface as loopstructural_b200.fault.FaultFrame, numerics from oracle/fault_oracle.py (test stand-in)."""

    calls = []

    def __init__(self, engines, world_to_local=None):
        from oracle import fault_oracle as fa

        self.fa = fa
        self.fields = [fa.FrameField(e.grid, e.c, None if world_to_local is None else np.asarray(world_to_local[i]), e.kind)
                       for i, e in enumerate(engines)]

    def _profile(self, p):
        fa = self.fa
        if not p.has_function:
            return fa.FaultProfile(None, 0.0)

        def one(q):
            return fa.Profile1D(int(q.kind), tuple(q.w), tuple(q.lim), float(q.min_v), float(q.max_v))

        return fa.FaultProfile(tuple(fa.AxisProfile(one(p.pos[a]), one(p.neg[a]) if p.composite[a] else None) for a in range(3)),
                               float(p.scale))

    def apply_to_points(self, xyz, profile, displacement, steps):
        OracleFaultFrame.calls.append("apply")
        return self.fa.apply_to_points(self.fields, self._profile(profile), displacement, steps, xyz)

    def evaluate_displacement(self, xyz, profile, displacement, fault_offset=0.0):
        OracleFaultFrame.calls.append("displacement")
        return self.fa.evaluate_displacement(self.fields, self._profile(profile), displacement, fault_offset, xyz)


eval_calls = []


def oracle_faulted_evaluate(frames, profiles, displacements, steps, engine, xyz, gradient=False, world_to_local=None):
    """Stand-in for loopstructural_b200.fault.faulted_evaluate (fault chain, then the field at the moved points)."""
    from oracle import fault_oracle as fa

    eval_calls.append(len(frames))
    pts = np.array(xyz, dtype=float)
    for fr, pr, d, s in zip(frames, profiles, displacements, steps):
        pts = fa.apply_to_points(fr.fields, fr._profile(pr), d, s, pts)
    fld = fa.FrameField(engine.grid, engine.c, world_to_local, engine.kind)
    return fld.gradient(pts) if gradient else fld.value(pts)


def _fault_model(ls, **kw):
    from LoopStructural.datasets import load_intrusion

    data, bb = load_intrusion()
    model = ls.GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    model.create_and_add_fault("fault", 500, nelements=2000, interpolatortype="FDI", buffer=0.3, **kw)
    model.create_and_add_foliation("strati", nelements=2000, interpolatortype="FDI", **kw)
    model.update()
    return model, bb


def test_fault_frame_and_faulted_foliation_through_the_model(ls, monkeypatch):
    """BASELINE configs[3] in small: a fault frame (three coupled FDI solves with orthogonality rows) and a foliation
    faulted by it, all with solver='b200'; FaultSegment.apply_to_points / evaluate_displacement go to the frame."""
    from loopstructural_b200 import fault as fl, plugin

    monkeypatch.setattr(fl, "FaultFrame", OracleFaultFrame)
    monkeypatch.setattr(fl, "faulted_evaluate", oracle_faulted_evaluate)
    OracleFaultFrame.calls.clear()
    model, bb = _fault_model(ls, solver="b200")
    eval_calls.clear()
    fault = model["fault"]
    for i in range(3):
        it = fault[i].interpolator
        assert type(it).__name__ == "B200FiniteDifferenceInterpolator" and it.latest_solve_timing["backend"] == "b200"
    assert model["strati"].interpolator.latest_solve_timing["backend"] == "b200"
    # the orthogonality families of coordinates 1 and 2 never existed on the host: generated from the resident fields
    assert [getattr(fault[i].interpolator._b200_engine, "device_families", 0) for i in range(3)] == [0, 2, 1]
    assert not any("orthogonal" in k for i in range(3) for k in fault[i].interpolator.constraints)
    assert "apply" in OracleFaultFrame.calls  # the foliation's data were un-faulted through the frame during set-up
    rng = np.random.default_rng(5)
    pts = bb[0] + rng.uniform(-0.2, 1.2, (400, 3)) * (bb[1] - bb[0])
    OracleFaultFrame.calls.clear()
    moved = fault.apply_to_points(pts)
    disp = fault.evaluate_displacement(pts)
    strati = model.evaluate_feature_value("strati", pts)
    assert OracleFaultFrame.calls.count("apply") >= 1 and "displacement" in OracleFaultFrame.calls
    assert 1 in eval_calls  # the faulted foliation: fault + field in one call (0 = plain evaluations with device projection)
    c_frame = [np.array(fault[i].interpolator.c) for i in range(3)]

    # an unsupported profile goes back to the stock method
    OracleFaultFrame.calls.clear()
    keep = fault._faultfunction
    fault._faultfunction = lambda gx, gy, gz: np.ones_like(gx)
    fault.apply_to_points(pts[:5])
    assert OracleFaultFrame.calls == []
    fault._faultfunction = keep

    plugin.uninstall()
    stock, _ = _fault_model(ls)
    sf = stock["fault"]
    assert sf.apply_to_points.__func__.__qualname__.startswith("FaultSegment.")
    assert type(sf[1].builder).install_gradient_constraint.__qualname__ == "GeologicalFeatureBuilder.install_gradient_constraint"
    assert any("orthogonal" in k for k in sf[1].interpolator.constraints)  # the stock build has the families on the host
    for i in range(3):
        cs = np.asarray(sf[i].interpolator.c)
        assert np.linalg.norm(c_frame[i] - cs) / np.sqrt(cs.size) / float(cs.max() - cs.min()) <= 1e-6  # the bar of BASELINE.json's north star
    ext = float(np.max(bb[1] - bb[0]))
    np.testing.assert_allclose(moved, sf.apply_to_points(pts), rtol=0, atol=1e-6 * ext)
    np.testing.assert_allclose(disp, sf.evaluate_displacement(pts), rtol=0, atol=1e-6 * abs(sf.displacement))
    s_ref = stock.evaluate_feature_value("strati", pts)
    np.testing.assert_array_equal(np.isnan(strati), np.isnan(s_ref))
    np.testing.assert_allclose(strati, s_ref, rtol=0, atol=1e-6 * float(np.nanmax(s_ref) - np.nanmin(s_ref)), equal_nan=True)


def test_p1_fault_frame_orthogonality_rows_come_from_the_device(ls):
    """A fault frame on tet meshes (interpolatortype='PLI', data shifted to a zero origin so that the barycentres stay inside
    the other coordinates' meshes): the '<feature>_orthogonal' families of coordinates 1 and 2 are generated by the engine
    (the P1 twin of lsb200_add_field_orthogonal_rows) and the solved frame matches the stock build."""
    from LoopStructural.datasets import load_intrusion
    from loopstructural_b200 import plugin

    def build(**kw):
        data, bb = load_intrusion()
        data = data.copy()
        data[["X", "Y", "Z"]] -= bb[0]
        bb = bb - bb[0]
        model = ls.GeologicalModel(bb[0, :], bb[1, :])
        model.data = data
        model.create_and_add_fault("fault", 500, nelements=2000, interpolatortype="PLI", buffer=0.3, **kw)
        model.update()
        return model

    fault = build(solver="b200")["fault"]
    for i in range(3):
        it = fault[i].interpolator
        assert type(it).__name__ == "B200P1Interpolator" and it.latest_solve_timing["backend"] == "b200"
    assert [getattr(fault[i].interpolator._b200_engine, "device_families", 0) for i in range(3)] == [0, 2, 1]
    assert not any("orthogonal" in k for i in range(3) for k in fault[i].interpolator.constraints)
    c_frame = [np.array(fault[i].interpolator.c) for i in range(3)]
    plugin.uninstall()
    sf = build()["fault"]
    assert any("orthogonal" in k for k in sf[1].interpolator.constraints)
    for i in range(3):
        cs = np.asarray(sf[i].interpolator.c)
        assert np.linalg.norm(c_frame[i] - cs) / np.sqrt(cs.size) / float(cs.max() - cs.min()) <= 1e-6


def _fold_model(ls, **kw):
    from loop_common.math import rng
    from LoopStructural.datasets import load_laurent2016

    # DiscreteFoldInterpolator shuffles the elements it puts fold constraints on with this module-level, unseeded
    # generator (LI/_discrete_fold_interpolator.py:116-176): same state for the two builds that are compared
    rng.bit_generator.state = np.random.default_rng(123).bit_generator.state
    data, bb = load_laurent2016()
    model = ls.GeologicalModel(bb[0, :], bb[1, :])
    model.data = data
    s2 = model.create_and_add_fold_frame("s2", nelements=1500, buffer=0.3, **kw)
    s1 = model.create_and_add_folded_fold_frame("s1", fold_frame=s2, limb_wl=4, av_fold_axis=True, nelements=1500, buffer=0.3, **kw)
    model.create_and_add_folded_foliation("s0", fold_frame=s1, limb_wl=1.0, av_fold_axis=True, nelements=1500, buffer=0.3, **kw)
    model.update()
    return model, bb


def test_refolded_model_through_the_b200_solver_matches_the_stock_solver(ls):
    """BASELINE configs[3], fold half (load_laurent2016 style): a fold frame (FDI), a folded fold frame and a folded
    foliation (DiscreteFoldInterpolator on a TetMesh) all solved with solver='b200'; the fold orientation / axis /
    normalisation rows and the three direction-weighted face families never exist on the host: the engine generates them
    from barycentres / faces and the fold vectors (here the oracle stand-in does, with oracle/p1_oracle.py)."""
    from loopstructural_b200 import plugin

    model, bb = _fold_model(ls, solver="b200")
    it0 = model["s0"].interpolator
    assert type(it0).__name__ == "B200DiscreteFoldInterpolator" and it0.latest_solve_timing["backend"] == "b200"
    fold_names = ["fold orientation", "fold axis", "fold normalisation", "fold regularisation 1", "fold regularisation 2",
                  "fold regularisation 3"]
    assert not set(fold_names) & set(it0.constraints)
    assert [f["name"] for f in it0._b200_generated] == fold_names
    assert [f["type"] for f in it0._b200_generated] == ["element"] * 3 + ["face"] * 3
    assert model["s1"][0].interpolator.latest_solve_timing["backend"] == "b200"   # DFI coordinate of the folded frame
    assert model["s2"][0].interpolator.latest_solve_timing["backend"] == "b200"   # FDI
    # Every stage against the stock solver ON THE SAME SYSTEM: the fold rows of s1 / s0 are functions of the previous
    # field (rotation angles fitted to its gradient) and the systems take thousands of CG iterations, so two separately
    # built models drift apart by far more than any solver difference -- the reference against itself under a one-ulp
    # change of s2 does the same.  Asking for a stock solver materialises the recorded families on the host (the same
    # elements, the same vectors) at the places the stock set-up would have put them.
    for it in (model["s1"][0].interpolator, it0):
        assert it._b200_reg == {} and it.interpolation_weights["cgw"] == 0.0
        # (s1: the fold vectors of four families hold NaNs, the reference drops such a family and so does the plug-in)
        generated = [f["name"] for f in it._b200_generated]
        assert len(generated) >= 2 and set(generated) <= set(fold_names) and not set(generated) & set(it.constraints)
        # the rows generated from barycentres / faces equal the host rows to rounding, not bit for bit, and these systems
        # take thousands of iterations: compared where CG has converged, not at the model's default tolerance
        tight = {"rtol": 1e-13, "atol": 0.0, "maxiter": 200000}
        assert it.solve_system("b200", solver_kwargs=tight) and it.latest_solve_timing["backend"] == "b200"
        c_b200 = np.array(it.c)
        assert it.solve_system("cg", solver_kwargs=tight) is True and it.latest_solve_timing["backend"] == "python"
        assert it._b200_generated == [] and list(it.constraints)[-len(generated):] == generated
        rng_ = float(it.c.max() - it.c.min())
        assert np.linalg.norm(c_b200 - it.c) / np.sqrt(it.c.size) / rng_ <= 1e-10
    c_frame = np.array(model["s2"][0].interpolator.c)

    plugin.uninstall()
    stock, _ = _fold_model(ls)
    its = stock["s0"].interpolator
    assert type(its).__name__ == "DiscreteFoldInterpolator" and its.latest_solve_timing["backend"] == "python"
    cs2 = np.asarray(stock["s2"][0].interpolator.c)
    assert np.linalg.norm(c_frame - cs2) / np.sqrt(cs2.size) / float(cs2.max() - cs2.min()) <= 1e-9  # the unfolded frame


def test_p1_directional_regularisation_rows_come_from_the_device(ls):
    """P1Interpolator.setup_interpolator(directional_regularisation=...) (LI/_p1interpolator.py:198-229,
    LI/_discrete_interpolator.py:413-451): the direction-weighted face-jump family is recorded with its vectors at the
    shared-face centres and generated by the engine; a vector field holding a NaN stays with the stock method, which
    drops the family; a stock solver afterwards sees the stock family."""
    import loop_interpolation as li
    from loop_common.supports import TetMesh

    mesh = TetMesh(origin=np.zeros(3), nsteps_cells=np.array([6, 5, 7]), step_vector=np.array([1.0, 1.2, 0.9]))
    rng = np.random.default_rng(11)
    pts = rng.uniform(0.3, 4.5, (60, 3))
    value = np.column_stack([pts, np.sin(pts[:, 0]) + 0.3 * pts[:, 2], np.ones(60)])

    def direction(x):
        x = np.asarray(x, float)
        return np.column_stack([np.cos(0.4 * x[:, 1]), 0.3 + 0.1 * x[:, 0], np.sin(0.5 * x[:, 2]) + 1.2])

    def with_nan(x):
        v = direction(x)
        v[3, 1] = np.nan
        return v

    terms = [{"weight": 0.4, "direction": direction, "name": "along"}, {"weight": 0.2, "direction": with_nan, "name": "broken"}]
    it = li.interpolator_map[li.InterpolatorType.PIECEWISE_LINEAR](mesh)
    assert type(it).__name__ == "B200P1Interpolator"
    it.set_value_constraints(value.copy())
    it.setup_interpolator(solver="b200", regularisation=0.05, directional_regularisation=terms)
    assert [(f["type"], f["name"]) for f in it._b200_generated] == [("face", "along")]
    assert "along" not in it.constraints and "broken" not in it.constraints and it._b200_reg == {"cgw": 0.05}
    it.reset()  # the stock reset empties the constraint dict: the recorded families go with it
    assert it._b200_generated == [] and it._b200_reg == {} and not it.constraints
    it.setup_interpolator(solver="b200", regularisation=0.05, directional_regularisation=terms)
    assert [(f["type"], f["name"]) for f in it._b200_generated] == [("face", "along")]
    tight = {"rtol": 1e-13, "atol": 0.0}
    assert it.solve_system("b200", solver_kwargs=tight) and it.latest_solve_timing["backend"] == "b200"
    assert it._b200_engine.device_families == 1
    c_b200 = np.array(it.c)
    assert it.solve_system("cg", solver_kwargs=tight) and it.latest_solve_timing["backend"] == "python"
    assert list(it.constraints)[:2] == ["edge jump", "along"] and it._b200_generated == []
    assert np.linalg.norm(c_b200 - it.c) / np.sqrt(it.c.size) / float(it.c.max() - it.c.min()) <= 1e-10


def test_fdi_directional_regularisation_rows_come_from_the_device(ls):
    """FiniteDifferenceInterpolator.setup_interpolator(directional_regularisation=...): the six direction-weighted
    second-derivative families (LI/_finite_difference_interpolator.py:1207-1340) are one record of the direction field,
    generated by the engine; a stock solver afterwards sees the six stock families."""
    import loop_interpolation as li
    from loop_common.supports import StructuredGrid

    sg = StructuredGrid(origin=np.zeros(3), nsteps_cells=np.array([7, 6, 8]), step_vector=np.array([1.0, 1.2, 0.9]))
    rng = np.random.default_rng(13)
    pts = rng.uniform(0.3, 5.5, (80, 3))
    value = np.column_stack([pts, np.sin(pts[:, 0]) + 0.3 * pts[:, 2], np.ones(80)])

    def direction(x):
        x = np.asarray(x, float)
        v = np.column_stack([np.cos(0.4 * x[:, 1]), 0.3 + 0.1 * x[:, 0], np.sin(0.5 * x[:, 2]) + 1.2])
        v[x[:, 0] > 4.0, 2] = 0.0
        return v

    it = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE](sg)
    assert type(it).__name__ == "B200FiniteDifferenceInterpolator"
    it.set_value_constraints(value.copy())
    it.setup_interpolator(solver="b200", regularisation=0.3, directional_regularisation=[{"weight": 0.8, "direction": direction, "name": "along"}])
    assert [(f["type"], f["name"]) for f in it._b200_generated] == [("directional", "along")]
    assert not any(k.startswith("along") for k in it.constraints)
    tight = {"rtol": 1e-13, "atol": 0.0}
    assert it.solve_system("b200", solver_kwargs=tight) and it.latest_solve_timing["backend"] == "b200"
    assert it._b200_engine.device_families == 1
    c_b200 = np.array(it.c)
    assert it.solve_system("cg", solver_kwargs=tight) and it.latest_solve_timing["backend"] == "python"
    assert sorted(k for k in it.constraints if k.startswith("along")) == sorted(f"along_{op}" for op in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz"))
    assert np.linalg.norm(c_b200 - it.c) / np.sqrt(it.c.size) / float(it.c.max() - it.c.min()) <= 1e-10


class _AnalyticFold:
    """Stands in for FoldEvent: three smooth vector fields (deformed orientation, fold axis, fold normal)."""

    def get_deformed_orientation(self, points):
        x = np.asarray(points, float)
        f = lambda a, b, c: np.column_stack([np.sin(a * x[:, 1]) + 0.3, np.cos(b * x[:, 2]) - 0.2 * x[:, 0], 0.5 + np.sin(c * x[:, 0])])  # noqa: E731
        return f(0.4, 0.7, 0.9), f(0.8, 0.3, 0.5), f(0.6, 0.5, 0.7)


@pytest.mark.parametrize("case", ["masked", "no_axis_step3", "nan_in_a_vector", "norm_sign_corrected"])
def test_fold_constraints_of_a_bare_fold_interpolator(ls, case):
    """DiscreteFoldInterpolator.add_fold_constraints through the plug-in without a model around it: a mask function
    (masked elements carry weight 0 in the reference and are left out here; masked faces get a zero vector), a family
    switched off and another element step, a fold vector holding a NaN (the reference drops that family, the plug-in hands it
    to the stock method which does) -- each against the stock class on the same generator state."""
    import loop_interpolation as li
    from loop_common.math import rng
    from loop_common.supports import TetMesh

    from loopstructural_b200 import plugin

    fold = _AnalyticFold()
    if case == "nan_in_a_vector":
        base = fold.get_deformed_orientation

        def with_nan(points):
            d, a, n = base(points)
            a = a.copy()
            a[::3, 2] = np.nan
            return d, a, n
        fold.get_deformed_orientation = with_nan
    kw = dict(fold_orientation=4.0, fold_axis_w=3.0, fold_normalisation=1.5, fold_norm=0.8, fold_regularisation=[0.2, 0.05, 0.03])
    if case == "masked":
        kw["mask_fn"] = lambda x: np.asarray(x)[:, 0] > 3.0
    if case == "no_axis_step3":
        kw.update(fold_axis_w=None, step=3)
    if case == "norm_sign_corrected":
        kw.update(dgz_alignment="correct")  # resolve_fold_norm_target flips the target when dgz opposes the normals
    pts = np.random.default_rng(21).uniform(0.3, 4.5, (50, 3))
    value = np.column_stack([pts, 0.4 * pts[:, 0] + np.cos(pts[:, 1]), np.ones(50)])
    tight = {"rtol": 1e-13, "atol": 0.0}

    def solve(solver):
        mesh = TetMesh(origin=np.zeros(3), nsteps_cells=np.array([6, 5, 6]), step_vector=np.array([1.0, 1.1, 0.9]))
        it = li.interpolator_map[li.InterpolatorType.DISCRETE_FOLD](mesh, fold)
        it.set_value_constraints(value.copy())
        if case == "norm_sign_corrected":
            npts = pts[:20]
            it.set_normal_constraints(np.column_stack([npts, -fold.get_deformed_orientation(npts)[2], np.ones(20)]))
        rng.bit_generator.state = np.random.default_rng(77).bit_generator.state
        it.setup_interpolator(solver=solver, regularisation=0.05, fold_weights=dict(kw)) if solver else \
            it.setup_interpolator(regularisation=0.05, fold_weights=dict(kw))
        state_after = rng.bit_generator.state["state"]["state"]
        assert it.solve_system(solver or "cg", solver_kwargs=tight)
        return it, state_after

    it, state_b200 = solve("b200")
    assert type(it).__name__ == "B200DiscreteFoldInterpolator" and it.latest_solve_timing["backend"] == "b200"
    names = [f["name"] for f in it._b200_generated]
    if case == "nan_in_a_vector":
        assert "fold axis" not in names and "fold axis" not in it.constraints and "fold regularisation 3" not in names
        assert names == ["fold orientation", "fold normalisation", "fold regularisation 1", "fold regularisation 2"]
    elif case == "no_axis_step3":
        assert names == ["fold orientation", "fold normalisation", "fold regularisation 1", "fold regularisation 2", "fold regularisation 3"]
        assert it._b200_generated[0]["data"].shape[0] == -(-it.support.n_elements // 3)
    elif case == "norm_sign_corrected":
        fam = [f for f in it._b200_generated if f["name"] == "fold normalisation"][0]
        assert np.all(fam["data"][:, 6] == -0.8 * 1.5)  # target = -fold_norm (corrected) * fold_normalisation
    else:
        assert len(names) == 6 and it._b200_generated[0]["data"].shape[0] < -(-it.support.n_elements // 2)
    c_b200 = np.array(it.c)
    plugin.uninstall()
    stock, state_stock = solve(None)
    assert type(stock).__name__ == "DiscreteFoldInterpolator"
    assert state_b200 == state_stock  # the generator was advanced exactly as the stock method advances it
    assert np.linalg.norm(c_b200 - stock.c) / np.sqrt(stock.c.size) / float(stock.c.max() - stock.c.min()) <= 1e-9


@pytest.mark.parametrize("case", ["weight_scale", "custom_operators"])
def test_set_ups_the_engine_cannot_apply_matrix_free_are_solved_from_explicit_rows(ls, case):
    """use_regularisation_weight_scale=True (per-node weights on the stencil rows, LI/_finite_difference_interpolator.py:643-692)
    and a caller-supplied operators= dict: nothing is diverted, the reference assembles its families, and solver='b200'
    solves exactly that system from explicit rows (families that ARE default masks with one weight still go matrix-free)."""
    import loop_interpolation as li
    from loop_common.supports import StructuredGrid
    from loop_interpolation._operator import Operator

    from loopstructural_b200 import plugin
    from oracle import fdi_oracle as fo

    grid, value, gradient = fo.synthetic_problem(12, 200, 30, seed=3)
    sg = StructuredGrid(origin=grid.origin, nsteps_cells=grid.nsteps - 1, step_vector=grid.step)
    kw = dict(use_regularisation_weight_scale=True) if case == "weight_scale" else \
        dict(operators={"laplacian": (Operator.Lapacian.astype(float), 0.7), "dxx": (Operator.Dxx_mask, 0.4)})
    tight = {"rtol": 1e-12, "atol": 0.0}

    def run(cls, solver):
        it = cls(sg)
        it.set_value_constraints(value.copy())
        it.set_normal_constraints(gradient.copy())
        it.setup_interpolator(**(dict(kw, solver=solver) if solver == "b200" else kw))
        assert it.solve_system(solver, solver_kwargs=tight)
        return it

    a = run(li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE], "b200")
    assert type(a).__name__ == "B200FiniteDifferenceInterpolator" and a.latest_solve_timing["backend"] == "b200"
    assert a._b200_generated == [] and ("laplacian" in a.constraints or case == "weight_scale")
    b = run(plugin._installed["stock"][li.InterpolatorType.FINITE_DIFFERENCE], "cg")
    assert b.latest_solve_timing["backend"] == "python"
    assert np.abs(a.c - b.c).max() <= 1e-9 * np.abs(b.c).max()


def test_two_faults_are_applied_as_one_chain(ls, monkeypatch):
    """A feature cut by two faults: BaseFeature._apply_faults hands both frames to lsb200_fault_displace_chain in one
    call (here an oracle stand-in composing the two displacements) and the model's answers equal the stock solver's."""
    from LoopStructural.datasets import load_intrusion

    from loopstructural_b200 import fault as fl, plugin

    chain_calls = []

    def oracle_chain(frames, profiles, displacements, steps, xyz):
        chain_calls.append(len(frames))
        pts = np.array(xyz, dtype=float)
        for fr, pr, d, s in zip(frames, profiles, displacements, steps):
            pts = fr.apply_to_points(pts, pr, d, s)
        return pts

    monkeypatch.setattr(fl, "FaultFrame", OracleFaultFrame)
    monkeypatch.setattr(fl, "apply_fault_chain", oracle_chain)
    monkeypatch.setattr(fl, "faulted_evaluate", oracle_faulted_evaluate)
    eval_calls.clear()

    def build(**kw):
        data, bb = load_intrusion()
        second = data[data["feature_name"] == "fault"].copy()
        second["feature_name"] = "fault_b"
        second["X"] = second["X"] + 0.35 * (bb[1, 0] - bb[0, 0])
        import pandas as pd

        model = ls.GeologicalModel(bb[0, :], bb[1, :])
        model.data = pd.concat([data, second], ignore_index=True)
        model.create_and_add_fault("fault", 300, nelements=1500, interpolatortype="FDI", buffer=0.3, **kw)
        model.create_and_add_fault("fault_b", -200, nelements=1500, interpolatortype="FDI", buffer=0.3, **kw)
        model.create_and_add_foliation("strati", nelements=1500, interpolatortype="FDI", **kw)
        model.update()
        return model, bb

    model, bb = build(solver="b200")
    assert len(model["strati"].faults) == 2
    rng = np.random.default_rng(6)
    pts = bb[0] + rng.uniform(0, 1, (300, 3)) * (bb[1] - bb[0])
    val = model.evaluate_feature_value("strati", pts)
    # evaluate_value of the faulted foliation: fault chain + field in ONE call; the chain alone serves _apply_faults
    assert 2 in eval_calls
    moved = model["strati"]._apply_faults(pts)
    assert chain_calls == [2] and np.abs(moved - pts).max() > 1.0
    plugin.uninstall()
    stock, _ = build()
    assert type(stock["strati"])._apply_faults.__qualname__ == "BaseFeature._apply_faults"
    assert type(stock["strati"]).evaluate_value.__qualname__ == "GeologicalFeature.evaluate_value"
    ref = stock.evaluate_feature_value("strati", pts)
    np.testing.assert_array_equal(np.isnan(val), np.isnan(ref))
    # seven coupled solves (two fault frames, then the foliation on data moved through both), each stopped at the
    # default tolerance: the fields agree to a few 1e-6 of the range
    np.testing.assert_allclose(val, ref, rtol=0, atol=2e-5 * float(np.nanmax(ref) - np.nanmin(ref)), equal_nan=True)


# ---- set-up / solver bookkeeping of the plug-in classes ----------------------------------------------
def _raw_fdi(ls, n=9, seed=2):
    from loop_common.supports import StructuredGrid
    from oracle import fdi_oracle as fo

    grid, value, gradient = fo.synthetic_problem(n, 80, 16, seed=seed)
    sg = StructuredGrid(origin=grid.origin, nsteps_cells=grid.nsteps - 1, step_vector=grid.step)
    return grid, sg, value, gradient


def _orth(grid, seed=7, n=25):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.05, 0.95, (n, 3)) * (grid.nsteps - 1) * grid.step + grid.origin
    vec = rng.normal(size=(n, 3))
    return pts, vec


def test_constraints_added_after_the_setup_survive_either_solver_direction(ls):
    """InterpolatorBuilder-style flow: setup WITHOUT solver='b200', constraints added afterwards, solve_system('b200') --
    and the opposite direction (diverted set-up, stock solver).  Neither re-runs the set-up; both equal the stock class."""
    import loop_interpolation as li

    grid, sg, value, gradient = _raw_fdi(ls)
    pts, vec = _orth(grid)

    def prepare(cls, **kw):
        it = cls(sg)
        it.set_value_constraints(value)
        it.set_normal_constraints(gradient)
        it.setup_interpolator(**kw)
        it.add_gradient_orthogonal_constraints(pts.copy(), vec.copy(), w=0.7, b=0.0, name="after_setup")  # it normalises vec in place
        return it

    cls = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE]
    assert cls.__name__ == "B200FiniteDifferenceInterpolator"
    stock = prepare(li.FiniteDifferenceInterpolator)
    assert stock.solve_system("cg", tol=1e-12)

    # explicit set-up -> b200: every family is in self.constraints; the twelve stencil families reach the engine as weights
    it = prepare(cls)
    assert not it._b200_divert and "dxx" in it.constraints and "after_setup" in it.constraints
    n_before = len(OracleEngine.instances)
    assert it.solve_system("b200", tol=1e-12) and it.latest_solve_timing["backend"] == "b200"
    eng = OracleEngine.instances[-1]
    assert len(OracleEngine.instances) == n_before + 1
    assert eng.weights == {k: 1.0 for k in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz", "dx", "dy", "dz")}
    assert len(eng.rows) == 5  # norm_0..2, value, after_setup: none of the 12 stencil families travelled as rows
    assert "after_setup" in it.constraints and "dxx" in it.constraints  # nothing was reset
    rng_ = float(stock.c.max() - stock.c.min())
    assert np.linalg.norm(it.c - stock.c) / np.sqrt(stock.c.size) / rng_ <= 1e-7  # family order differs: other rounding
    # a second solve reuses the engine (seven node vectors of device memory) instead of re-creating it
    assert it.solve_system("b200", tol=1e-12) and len(OracleEngine.instances) == n_before + 1 and len(eng.rows) == 5
    # c is an ordinary ndarray that owns its memory
    assert it.c.flags.owndata or it.c.base is None or isinstance(it.c.base, np.ndarray)

    # diverted set-up -> stock solver: the recorded families are assembled where the stock set-up puts them
    it2 = prepare(cls, solver="b200")
    assert it2._b200_divert and "dxx" not in it2.constraints and "after_setup" in it2.constraints
    assert it2.solve_system("cg", tol=1e-12) and it2.latest_solve_timing["backend"] == "python"
    assert list(it2.constraints) == list(stock.constraints)
    np.testing.assert_array_equal(it2.c, stock.c)  # same matrix, same order: the same bits


def test_rectilinear_grids_are_refused_by_the_b200_solver(ls):
    import loop_interpolation as li
    from loop_common.supports import RectilinearGrid

    cls = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE]
    x = np.array([0.0, 0.1, 0.3, 0.6, 1.0])
    rg = RectilinearGrid(x, x, x)
    it = cls(rg)
    rng = np.random.default_rng(0)
    it.set_value_constraints(np.column_stack([rng.uniform(0.1, 0.9, (20, 3)), rng.normal(size=20), np.ones(20)]))
    it.setup_interpolator(solver="b200")
    assert not it._b200_divert
    with pytest.raises(ValueError, match="regular StructuredGrid"):
        it.solve_system("b200")
    assert it.solve_system("cg") and not OracleEngine.instances


def test_diverted_setup_never_touches_the_node_table(ls, monkeypatch):
    """with solver='b200' neither support.nodes (24 B/node) nor the all-node KD-tree of
    add_norm_constraints is built; the norm families are still the stock ones bit for bit."""
    import loop_interpolation as li
    from loop_common.supports import StructuredGrid

    grid, sg, value, gradient = _raw_fdi(ls, n=10, seed=5)
    stock = li.FiniteDifferenceInterpolator(sg)
    stock.set_value_constraints(value)
    stock.set_normal_constraints(gradient)
    stock.setup_interpolator()
    monkeypatch.setattr(StructuredGrid, "nodes", property(lambda self: (_ for _ in ()).throw(AssertionError("support.nodes built"))))
    cls = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE]
    it = cls(sg)
    it.set_value_constraints(value)
    it.set_normal_constraints(gradient)
    it.setup_interpolator(solver="b200")
    assert it.regularisation_scale.strides == (0,)  # a broadcast view, not dof doubles
    for k in ("norm_0", "norm_1", "norm_2", "value"):
        assert (it.constraints[k]["matrix"] != stock.constraints[k]["matrix"]).nnz == 0
        np.testing.assert_array_equal(it.constraints[k]["b"], stock.constraints[k]["b"])
        np.testing.assert_array_equal(it.constraints[k]["w"], stock.constraints[k]["w"])
    assert it.solve_system("b200")


def _inequality_problem(ls, n=12):
    import loop_interpolation as li
    from loop_common.supports import StructuredGrid
    from oracle import fdi_oracle as fo

    grid, value, gradient = fo.synthetic_problem(n, 90, 15, seed=3)
    sg = StructuredGrid(origin=grid.origin, nsteps_cells=grid.nsteps - 1, step_vector=grid.step)
    rng = np.random.default_rng(0)
    pts = rng.uniform(0.1, 0.9, (25, 3))
    f = fo.synthetic_field(pts)
    ineq = np.column_stack([pts, f + 0.05, f + 0.3, np.ones(len(pts))])  # push the field above what the data say

    def make(cls, **kw):
        it = cls(sg)
        it.set_value_constraints(value.copy())
        it.set_gradient_constraints(gradient.copy())
        it.set_value_inequality_constraints(ineq.copy())
        it.setup_interpolator(**kw)
        return it

    return li, make, pts, f


def test_admm_inner_solves_run_on_the_engine_and_match_the_stock_admm(ls):
    """SURVEY 8 f rank 3: solver='admm' with linsys_solver='b200' keeps the reference's outer ADMM loop and sends every
    inner least-squares solve to the engine (matrix-free stencil + data rows + rho Q as a row block, new right-hand side
    per outer iteration).  With both inner solvers converged the two runs follow the same outer iterates."""
    li, make, pts, f = _inequality_problem(ls)
    cls = li.interpolator_map[li.InterpolatorType.FINITE_DIFFERENCE]
    kw = {"nmajor": 25, "rtol": 1e-13}
    stock = make(li.FiniteDifferenceInterpolator)
    assert stock.solve_system("admm", solver_kwargs=dict(kw, linsys_solver="cg"))
    for setup_kw in ({}, {"solver": "b200"}):  # explicit and diverted set-up
        it = make(cls, **setup_kw)
        n_eng = len(OracleEngine.instances)
        assert it.solve_system("admm", solver_kwargs=dict(kw, linsys_solver="b200"))
        t = it.latest_solve_timing
        assert t["inner_backend"] == "b200" and t["inner_solves"] == 25 and len(OracleEngine.instances) == n_eng + 1
        eng = OracleEngine.instances[-1]
        assert eng.weights == {k: 1.0 for k in ("dxx", "dyy", "dzz", "dxy", "dxz", "dyz", "dx", "dy", "dz")}
        assert len(eng.rows) == 4  # gradient, gradient_0, value and the rho Q block: no stencil family as rows
        rng_ = float(stock.c.max() - stock.c.min())
        assert np.linalg.norm(it.c - stock.c) / np.sqrt(stock.c.size) / rng_ <= 1e-7
        np.testing.assert_allclose(it.evaluate_value(pts), stock.evaluate_value(pts), rtol=0, atol=1e-7 * rng_)
