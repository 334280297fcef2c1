"""Only needed where the reference tree is absent; a deployment keeps backend/config.py."""
from monte_carlo_retirement_b200.config import (  # noqa: F401
    Config,
    ConfigurationError,
    OtherIncomeStreamConfig,
    load_config_from_json,
)
