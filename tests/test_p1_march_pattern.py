"""loopstructural_b200/csrc/p1_march_taps.inc (the baked tap pattern of k_p1_march) is what scripts/gen_p1_march_taps.py
produces from p1_tables.regularisation_table today, and every class has the tap counts the kernel's comments quote."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_pattern_is_current():
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_p1_march_taps.py"), "--check"], cwd=ROOT)
    assert rc.returncode == 0, "rerun scripts/gen_p1_march_taps.py"


def test_pattern_shape():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gen_p1_march_taps as gen

    pats = gen.pattern()
    assert [len(p) for p in pats] == [15, 33, 33, 15, 33, 15, 15, 33]
    for cls, p in enumerate(pats):
        assert (0, 0, 0) in p
        # symmetric operator: the mirror tap belongs to the class of the neighbour it points at
        for dx, dy, dz in p:
            other = ((cls & 1) + dx) % 2 + 2 * ((((cls >> 1) & 1) + dy) % 2) + 4 * (((cls >> 2) + dz) % 2)
            assert (-dx, -dy, -dz) in pats[other]
