"""Constants shared by make_golden.py (generator, needs the reference) and tests/conftest.py (loader, must not)."""
CUBIC_SETS = [("bezier", None), ("bspline", 7), ("catmull_rom", 4), ("hermite", 12)]   # (basis, tessellation rate; None = the default 4)
