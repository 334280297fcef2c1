"""Several GPUs from ONE process (monte_carlo_retirement_b200/multi_device.py): worker threads, path
shards of one Philox stream, all-reduces by hand-written kernels over NVLink peer memory
(csrc/mcr_comm.cu). Needs >= 2 devices; bit-for-bit equality with the single-GPU engine is the bar
(results must not depend on the device count)."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

import scenarios
from gpu_util import make_sim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (a spinning peer all-reduce must never share a GPU with its peer)")
    return list(range(n))


def _multi(cfg, devices, **kw):
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.multi_device import MultiDeviceSimulator

    return MultiDeviceSimulator(Config(**cfg), devices=devices, **kw)


def test_peer_all_reduce_kernels():
    """Every op / dtype of csrc/mcr_comm.cu against torch on the host, sizes from 1 element to several
    staging chunks, many calls back to back (the two staging halves are re-used every second call)."""
    import threading

    import torch

    from monte_carlo_retirement_b200.multi_device import PeerCollectives, PeerGroup

    devices = _devices()
    group = PeerGroup(devices)
    rng = np.random.default_rng(3)
    cases = []
    for n in (1, 7, 2048, 100_003, 5_000_000):
        for dtype, what in ((np.int32, "sum"), (np.int64, "sum"), (np.int64, "min"), (np.int64, "max"),
                            (np.float64, "sum"), (np.float64, "min"), (np.float64, "max")):
            if n == 5_000_000 and what != "sum":
                continue
            data = [(rng.integers(-1000, 1000, n).astype(dtype) if dtype != np.float64 else rng.normal(size=n))
                    for _ in devices]
            cases.append((what, data))
    results, errors = {}, []

    def worker(rank):
        try:
            torch.cuda.set_device(devices[rank])
            sim = make_sim(scenarios.TEST_BASE, device=devices[rank])
            coll = PeerCollectives(group, rank, sim.native_context)
            outs = []
            for what, data in cases:
                t = torch.from_numpy(data[rank]).to(f"cuda:{devices[rank]}")
                getattr(coll, what + "_")(t)
                outs.append(t.cpu().numpy())
            coll.barrier()
            results[rank] = outs
            coll.close()
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k, (what, data) in enumerate(cases):
        stack = np.stack(data)
        want = {"sum": stack.sum(0), "min": stack.min(0), "max": stack.max(0)}[what]
        for rank in range(len(devices)):
            got = results[rank][k]
            if what == "sum" and stack.dtype == np.float64:
                # the kernel adds rank 0, 1, 2, ... in order: the same association as this loop
                acc = data[0].copy()
                for d in data[1:]:
                    acc = acc + d
                want = acc
            assert np.array_equal(got, want), (what, stack.dtype, stack.shape, rank)


def test_multi_device_simulator_equals_single_device_bit_for_bit():
    devices = _devices()
    cfg = dict(scenarios.SYNTH_C3_VOL, num_simulations_search=100_000, target_probability=60.0)
    n, wm = 300_001, 150
    one = make_sim(cfg, device=0)
    many = _multi(cfg, devices)
    try:
        a, b = many.run_aggregates(wm, n, samples=True), one.run_aggregates(wm, n, samples=True)
        for k in b:
            va, vb = a[k], b[k]
            if hasattr(vb, "to_numpy"):
                assert np.array_equal(va.to_numpy(), vb.to_numpy(), equal_nan=True), k
            else:
                assert va == vb or (va != va and vb != vb), (k, va, vb)
        ta, tb = many.run_monte_carlo_simulations(wm, n), one.run_monte_carlo_simulations(wm, n)
        assert ta[0].equals(tb[0])                        # all N rows, global path order, zero-copy host block
        keep = ta[0].copy(deep=True)
        t2 = many.run_monte_carlo_simulations(wm + 12, n)  # a second call must not overwrite a live result
        assert ta[0].equals(keep) and not t2[0].equals(keep)
        for i in (1, 3, 4):
            assert np.array_equal(ta[i].to_numpy(), tb[i].to_numpy(), equal_nan=True), i
        assert ta[2] == tb[2] and ta[5] == tb[5] and ta[6] == tb[6]
        events_a, events_b = [], []
        ra = many.find_minimum_working_months(verbose=False, progress_callback=events_a.append)
        rb = one.find_minimum_working_months(verbose=False, progress_callback=events_b.append)
        assert ra == rb and events_a == events_b
        # small calls run on the first device alone: same numbers
        sa, sb = many.run_monte_carlo_simulations(wm, 500), one.run_monte_carlo_simulations(wm, 500)
        assert sa[0].equals(sb[0]) and sa[6] == sb[6]
    finally:
        many.close()


def test_dropin_uses_every_visible_device(monkeypatch):
    """`from simulation import RetirementMonteCarloSimulator` (dropin/) on a multi-GPU box: the class
    the reference's main.py / server.py construct drives all devices, no launcher involved."""
    devices = _devices()
    monkeypatch.syspath_prepend(os.path.join(ROOT, "dropin"))
    for name in ("simulation", "config", "constants"):
        sys.modules.pop(name, None)
    import simulation as dropin_simulation
    from config import Config

    from monte_carlo_retirement_b200.multi_device import MultiDeviceSimulator

    cfg = Config(**dict(scenarios.SYNTH_C3))
    sim = dropin_simulation.RetirementMonteCarloSimulator(cfg)
    try:
        assert isinstance(sim, MultiDeviceSimulator) and sim.devices == devices
        out = sim.run_monte_carlo_simulations(working_months=240, num_simulations=400_000)
        ref = make_sim(scenarios.SYNTH_C3, device=0).run_monte_carlo_simulations(240, 400_000)
        assert out[0].equals(ref[0]) and np.array_equal(out[1].to_numpy(), ref[1].to_numpy())
    finally:
        sim.close()
    monkeypatch.setenv("MCR_DEVICES", "0")
    single = dropin_simulation.RetirementMonteCarloSimulator(cfg)
    assert not isinstance(single, MultiDeviceSimulator)
    for name in ("simulation", "config", "constants"):
        sys.modules.pop(name, None)
