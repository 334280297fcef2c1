"""monte_carlo_retirement_b200 — B200-native (sm_100a CUDA) Monte Carlo retirement path engine,
a drop-in for the `simulate` entry points of rflamino/monte_carlo_retirement's
backend/simulation.py. See DESIGN.md and include/mcr.h."""
from .constants import MONTHS_PER_YEAR, SMALL_EPSILON  # noqa: F401

__all__ = ["MONTHS_PER_YEAR", "SMALL_EPSILON"]
__version__ = "0.1.0"
