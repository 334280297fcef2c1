"""Flat-module shim: put this directory on sys.path AHEAD of the reference's backend/ (or copy
this file over backend/simulation.py) and `from simulation import ...` in backend/main.py,
backend/server.py, backend/plotting.py and tests/test_simulation_correctness.py resolves to the
B200 engine with the reference's names."""
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator as _SingleDeviceSimulator
from monte_carlo_retirement_b200.simulation import (  # noqa: F401
    age_at_retirement_year,
    arithmetic_to_log_params,
    median_first_year_withdrawal_rate,
    retirement_age,
    stream_payment_start_age,
    stream_payment_start_month_index,
    trajectory_time_points,
    years_from_t0_to_age,
)
from monte_carlo_retirement_b200.constants import MONTHS_PER_YEAR, SMALL_EPSILON  # noqa: F401


class RetirementMonteCarloSimulator(_SingleDeviceSimulator):
    """The reference's class name. With more than one visible device (MCR_DEVICES = "all" — the
    default —, "0", "0,1,...") construction returns a MultiDeviceSimulator: one process, one worker
    thread per GPU, path shards of one Philox stream, all-reduces over NVLink peer memory; results
    are bit-identical for any device count and calls too small to shard run on one device. So
    backend/main.py and backend/server.py use every GPU of the box without any launcher."""

    def __new__(cls, params_model, main_seed_override=None, **kw):
        if cls is RetirementMonteCarloSimulator and "device" not in kw:
            from monte_carlo_retirement_b200.multi_device import MultiDeviceSimulator, devices_from_env

            devices = devices_from_env()
            if len(devices) > 1:
                return MultiDeviceSimulator(params_model, main_seed_override, devices=devices, **kw)
        return super().__new__(cls)
