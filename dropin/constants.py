from monte_carlo_retirement_b200.constants import *  # noqa: F401,F403
