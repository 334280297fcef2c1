"""pytest configuration: `gpu` marker + import paths.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks,
gloo world_size-2 tests. `-m gpu` runs on a B200 and calls the CUDA path through the C-ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # the suites bind the in-tree engine library: (re)build it when it is missing or stale (nvcc
    # cross-compiles sm_100a without a GPU) and the CPU oracle next to it
    from monte_carlo_retirement_b200 import build as mcr_build

    if mcr_build.needs_build():
        mcr_build.build()
    from oracle import oracle as orc

    orc.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
