"""Scenario types at the drop-in boundary.

The reference's Pydantic models (/root/reference/backend/config.py:12-126) are pure
validation and are out of scope for the GPU engine: a deployment keeps
using the reference's own `config.py`. This module exists so that the package is usable and
testable where the reference tree is absent (the GPU box): same field names, bounds, alias
(`scenario` -> `Nickname`) and derived `allocation_inv2_pct`, so a Config built here and one
built by the reference are interchangeable for `RetirementMonteCarloSimulator`.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional

from pydantic import BaseModel, Field, field_validator

try:  # the reference logs through loguru; fall back to logging when it is absent
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("mcr_b200")


class ConfigurationError(Exception):
    """The configuration file cannot be read or parsed (config.py:8-9)."""


def _unit(**kw):
    return Field(..., ge=0.0, le=1.0, **kw)


class OtherIncomeStreamConfig(BaseModel):
    """One age-gated income stream paid during retirement (config.py:12-47)."""

    name: str
    monthly_amount_today: float = Field(..., ge=0)
    start_at_age: float = Field(..., ge=0, le=120)
    duration_years: Optional[int] = Field(None, ge=0)
    inflation_indexed: bool = True
    tax_rate: float = _unit()


class Config(BaseModel):
    """Scenario (config.py:48-126). Field order follows the kernel's parameter block."""

    model_config = {"validate_by_name": True, "validate_assignment": True}

    Nickname: str = Field("DefaultScenario", alias="scenario")
    initial_balance: float = Field(..., ge=0)
    monthly_contribution: float = Field(..., ge=0)
    contribution_growth_rate_annual: float = Field(0.0, ge=0)
    monthly_expenses: float = Field(..., ge=0)
    current_age: float = Field(..., ge=0, le=120)
    retirement_years: int = Field(..., gt=0)

    allocation_inv1_pct: float = _unit()
    inv1_returns_mean: float = Field(..., gt=-1.0)
    inv1_returns_volatility: float = Field(..., ge=0.0)
    inv1_annual_tax_on_gains_rate: float = _unit()
    inv1_realized_gains_tax_rate: float = Field(0.0, ge=0.0, le=1.0)
    inv1_use_realized_gains_tax_system: bool = False

    inv2_premium_over_inflation_mean: float = Field(..., gt=-1.0)
    inv2_premium_over_inflation_volatility: float = Field(..., ge=0.0)
    inv2_annual_tax_on_gains_rate: float = _unit()
    inv2_realized_gains_tax_rate: float = Field(0.0, ge=0.0, le=1.0)
    inv2_use_realized_gains_tax_system: bool = True

    inflation_rate_mean: float = Field(..., gt=-1.0)
    inflation_rate_volatility: float = Field(..., ge=0.0)
    equity_inflation_correlation: float = Field(0.0, ge=-1.0, le=1.0)

    num_simulations_main: int = Field(..., gt=0)
    num_simulations_search: int = Field(..., gt=0)
    target_probability: float = Field(..., ge=0.0, le=100.0)
    starting_working_months_search: int = Field(..., ge=0)
    seed: Optional[int] = Field(None, ge=0)
    num_processes: Optional[int] = Field(1, ge=1)  # accepted, ignored: paths run on the GPU

    other_income_streams: List[OtherIncomeStreamConfig] = Field(default_factory=list)

    @field_validator("inflation_rate_volatility")
    @classmethod
    def _warn_high_inflation_vol(cls, v: float) -> float:
        if v > 0.05:
            logger.warning(f"Inflation volatility ({v * 100:.1f}%) is relatively high.")
        return v

    @field_validator("inv1_returns_volatility")
    @classmethod
    def _warn_low_equity_vol(cls, v: float) -> float:
        if v < 0.05:
            logger.warning(f"Equity (Inv1) volatility ({v * 100:.1f}%) is unusually low; "
                           "sequence-of-returns risk will be understated.")
        return v

    @property
    def allocation_inv2_pct(self) -> float:
        return 1.0 - self.allocation_inv1_pct


def load_config_from_json(file_path: str) -> Dict[str, Any]:
    """config.py:129-144."""
    if not os.path.exists(file_path):
        raise ConfigurationError(f"Configuration file not found at: {file_path}")
    try:
        with open(file_path, "r", encoding="utf-8") as fh:
            return json.load(fh)
    except json.JSONDecodeError as exc:
        raise ConfigurationError(f"Error parsing JSON file '{file_path}': {exc}") from exc
    except Exception as exc:
        raise ConfigurationError(f"Unexpected error reading config file '{file_path}': {exc}") from exc
