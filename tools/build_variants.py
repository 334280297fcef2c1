#!/usr/bin/env python
"""Build tuning variants of libmcr_b200.so side by side (each selected at run time with
MCR_LIB=<path>), then restore the default build. For A/B runs in ONE gpurun call, e.g.

    python tools/build_variants.py mb8:MCR_MIN_BLOCKS=8 mb4:MCR_MIN_BLOCKS=4 b256:MCR_BLOCK=256,MCR_MIN_BLOCKS=3
    gpurun -- 'for v in "" _mb8 _mb4 _b256; do MCR_LIB=$PWD/monte_carlo_retirement_b200/_lib/libmcr_b200$v.so \
        python tools/run_timeline.py; done'

The variant files are git-ignored (*.so) but travel with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "monte_carlo_retirement_b200", "_lib")
LIB = os.path.join(LIB_DIR, "libmcr_b200.so")


def build(env_extra):
    env = dict(os.environ)
    for k in ("MCR_MIN_BLOCKS", "MCR_BLOCK", "MCR_SCAN_UNROLL", "MCR_FIRST_CHUNK"):
        env.pop(k, None)
    env.update(env_extra)
    subprocess.run([sys.executable, "-m", "monte_carlo_retirement_b200.build", "--force"], cwd=ROOT, env=env, check=True,
                   stdout=subprocess.DEVNULL)


def main():
    for spec in sys.argv[1:]:
        name, _, macros = spec.partition(":")
        env_extra = dict(m.split("=", 1) for m in macros.split(",") if m)
        build(env_extra)
        out = os.path.join(LIB_DIR, f"libmcr_b200_{name}.so")
        shutil.copyfile(LIB, out)
        res = subprocess.run(["cuobjdump", "-res-usage", out], capture_output=True, text=True).stdout
        regs = [ln for ln in res.splitlines() if "REG:" in ln]
        print(f"{name}: {env_extra} -> {out} ({len(regs)} kernels, max REG "
              f"{max(int(ln.split('REG:')[1].split()[0]) for ln in regs)})")
    build({})  # the default build is what tests, bench.py and the driver load
    print("default build restored:", LIB)


if __name__ == "__main__":
    main()
