"""Import stub so the reference (which hard-imports pyvista) loads without the viewer stack.

Test infrastructure only: used by tests/golden/make_golden.py in the build container.
"""


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()


class CellType:
    VOXEL = 11
    TETRA = 10
    QUAD = 9
    TRIANGLE = 5
    HEXAHEDRON = 12


def __getattr__(name):
    return _Dummy()
