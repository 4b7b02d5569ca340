"""Import stub (see pyvista stub)."""
