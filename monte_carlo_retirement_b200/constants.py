"""Boundary constants (mirror of /root/reference/backend/constants.py:1-3); the CUDA side has
the same values in include/mcr.h (MCR_MONTHS_PER_YEAR, MCR_SMALL_EPSILON)."""
MONTHS_PER_YEAR: int = 12
SMALL_EPSILON: float = 1e-6
DEFAULT_PLOT_FILENAME: str = "retirement_projection.png"
