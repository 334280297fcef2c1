"""oracle/_ref — the UNMODIFIED reference, staged so that it can travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rules as the rest of oracle/: tests/,
__graft_entry__ and the CPU legs of bench.py). The reference is pure Python, so "building" it
means copying the five backend modules of the path and the reference's own test file from the
read-only checkout into the git-ignored oracle/_ref/ (never into the repository history):

    oracle/_ref/backend/{simulation,config,constants,utils,server}.py
    oracle/_ref/tests/test_simulation_correctness.py

Used by
  * bench.py (`cpu_baseline.kind == "reference"`, `--impl reference`): times
    RetirementMonteCarloSimulator.run_monte_carlo_simulations of the real reference with its own
    multiprocessing.Pool on the box's host cores (backend/simulation.py:952-1128, :996-1001);
  * tests/test_gpu_reference_suite.py: runs the reference's 23 tests, unmodified, against dropin/.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("MCR_REFERENCE_ROOT", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")
BACKEND_FILES = ("simulation.py", "config.py", "constants.py", "utils.py", "server.py")
TEST_FILES = ("test_simulation_correctness.py",)


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, "backend", f)) for f in BACKEND_FILES)


def build(force: bool = False) -> bool:
    """Stage the reference when its checkout is mounted (the build container); returns whether
    oracle/_ref is usable afterwards."""
    if not os.path.isdir(os.path.join(REF_SRC, "backend")):
        return available()
    for sub, names in (("backend", BACKEND_FILES), ("tests", TEST_FILES)):
        os.makedirs(os.path.join(REF_DIR, sub), exist_ok=True)
        for name in names:
            src, dst = os.path.join(REF_SRC, sub, name), os.path.join(REF_DIR, sub, name)
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copyfile(src, dst)
    for name in ("config.json", "jorge.json"):
        if os.path.exists(os.path.join(REF_SRC, name)):
            shutil.copyfile(os.path.join(REF_SRC, name), os.path.join(REF_DIR, name))
    return available()


def backend_path() -> str:
    return os.path.join(REF_DIR, "backend")


if __name__ == "__main__":
    print("oracle/_ref ready" if build(force=True) else "reference checkout not found; oracle/_ref not staged")
