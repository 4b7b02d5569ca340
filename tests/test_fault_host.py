"""Host logic of loopstructural_b200/fault.py (no device): struct layout against include/lsb200.h, translation of
the reference's profile objects, transform rows."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from loopstructural_b200 import _native as nat
from loopstructural_b200 import fault as fl

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def test_struct_layout_matches_the_header():
    # lsb200_profile1d: 2 x int32 + 8 doubles; lsb200_fault_profile: 4 x int32 + double + 6 profiles
    assert C.sizeof(nat.Profile1D) == 72 and nat.Profile1D.w.offset == 8 and nat.Profile1D.max_v.offset == 64
    assert C.sizeof(nat.FaultProfile) == 16 + 8 + 6 * 72
    assert nat.FaultProfile.scale.offset == 16 and nat.FaultProfile.pos.offset == 24 and nat.FaultProfile.neg.offset == 24 + 216


def test_transform_rows():
    w2l = np.eye(4)
    ang = 0.7
    w2l[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    w2l[:3, 3] = [3.0, -2.0, 4500.0]
    t = fl.transform_rows(w2l)
    assert t.shape == (21,)
    np.testing.assert_array_equal(t[:12].reshape(3, 4), w2l[:3])
    np.testing.assert_allclose(t[12:].reshape(3, 3), w2l[:3, :3].T, atol=1e-15)
    np.testing.assert_array_equal(fl.transform_rows(None), np.concatenate([np.eye(4)[:3].ravel(), np.eye(3).ravel()]))


def test_fault_profile_builder():
    cub = fl.profile1d(fl.PROFILE_CUBIC, (1, 2, 3, 4), (0, 1), -1, 1)
    p = fl.fault_profile([(cub, fl.profile1d(fl.PROFILE_ZEROS)), fl.profile1d(fl.PROFILE_ONES), cub], 0.25)
    assert p.has_function == 1 and list(p.composite) == [1, 0, 0] and p.scale == 0.25
    assert list(p.pos[0].w) == [1, 2, 3, 4] and p.neg[0].kind == fl.PROFILE_ZEROS and p.neg[2].kind == fl.PROFILE_CUBIC
    assert fl.fault_profile(None).has_function == 0
    with pytest.raises(ValueError):
        fl.fault_profile([cub, cub])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "packages")), reason="reference not mounted")
def test_profiles_of_the_reference_objects_translate_and_evaluate_alike():
    sys.dont_write_bytecode = True
    for p in (os.path.join(HERE, "golden", "_stubs"), f"{REF}/packages/loop_common/src",
              f"{REF}/packages/loop_interpolation/src", REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging

    logging.disable(logging.CRITICAL)
    from LoopStructural.modelling.features.fault import _fault_function as ff

    from oracle import fault_oracle as fa

    def oracle_of(p):
        def one(q):
            return fa.Profile1D(int(q.kind), tuple(q.w), tuple(q.lim), float(q.min_v), float(q.max_v))

        return fa.FaultProfile(tuple(fa.AxisProfile(one(p.pos[a]), one(p.neg[a]) if p.composite[a] else None) for a in range(3)),
                               float(p.scale))

    rng = np.random.default_rng(0)
    g = rng.uniform(-1.6, 1.6, (3, 4000))
    g[:, :7] = [[0.0, 1.0, -1.0, 0.5, -0.5, 1.0000001, -1.0000001]] * 3
    for fd in (ff.BaseFault().fault_displacement, ff.BaseFault3D().fault_displacement,
               ff.FaultDisplacement(hw=ff.BaseFault.hw, fw=ff.Zeros(), gy=ff.Ones(), gz=ff.Ones(), scale=2.0)):
        p = fl.profile_from_reference(fd)
        assert p is not None and p.has_function == 1 and p.composite[0] == 1
        np.testing.assert_allclose(oracle_of(p)(*g), fd(*g), rtol=1e-14, atol=1e-15)
    assert fl.profile_from_reference(None).has_function == 0
    # anything outside the closed family is refused, never approximated
    assert fl.profile_from_reference(lambda gx, gy, gz: gx) is None
    assert fl.profile_from_reference(ff.FaultDisplacement(gx=ff.Ones(), gy=np.tanh, gz=ff.Ones())) is None
    assert fl.profile_from_reference(ff.FaultDisplacement(hw=ff.Ones(), fw=np.tanh)) is None


def test_chain_argument_packing():
    """Host packing for lsb200_fault_displace_chain / lsb200_faulted_eval: 3 handles and 63 transform doubles per fault in
    fault order, identity rows for a frame without transforms when another one has them, placeholders for no faults."""
    class Frame:
        def __init__(self, base, transforms):
            self._handles = (nat.Handle * 3)(nat.Handle(base), nat.Handle(base + 1), nat.Handle(base + 2))
            self._transforms = transforms

    w2l = np.eye(4)
    w2l[:3, 3] = [1.0, 2.0, 3.0]
    t = np.concatenate([fl.transform_rows(w2l)] * 3)
    frames = [Frame(10, None), Frame(20, t)]
    profs = [fl.fault_profile(None), fl.fault_profile([fl.profile1d(fl.PROFILE_ONES)] * 3, 0.25)]
    nf, handles, tr, parr, disp, st = fl._chain_args(frames, profs, [1.5, -2.0], [4, 7])
    assert nf == 2 and [handles[i] for i in range(6)] == [10, 11, 12, 20, 21, 22]
    assert tr.shape == (126,) and np.array_equal(tr[63:], t)
    np.testing.assert_array_equal(tr[:21], fl.transform_rows(None))
    assert parr[0].has_function == 0 and parr[1].scale == 0.25 and list(disp) == [1.5, -2.0] and list(st) == [4, 7]
    nf, handles, tr, parr, disp, st = fl._chain_args([], [], [], [])
    assert nf == 0 and tr is None and len(handles) == 1 and len(parr) == 1 and len(st) == 1
    with pytest.raises(ValueError):
        fl._chain_args(frames, profs, [1.0], [4, 7])
