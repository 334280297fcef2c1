"""Loader for the golden fixtures written by tests/golden/make_golden.py."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Iterator, Tuple

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def path_fixture_names():
    return sorted(f[len("paths_"):-len(".npz")] for f in os.listdir(GOLDEN_DIR)
                  if f.startswith("paths_") and f.endswith(".npz"))


def load_paths(name: str) -> Tuple[Dict[str, Any], int, list]:
    z = np.load(os.path.join(GOLDEN_DIR, f"paths_{name}.npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg_json"]))
    cases = []
    for k in range(int(z["n_cases"])):
        pre = f"case{k}_"
        cases.append({key[len(pre):]: (z[key].item() if z[key].ndim == 0 else z[key])
                      for key in z.files if key.startswith(pre)})
    return cfg, int(z["main_seed"]), cases


def iter_cases() -> Iterator[Tuple[str, Dict[str, Any], int, Dict[str, Any]]]:
    for name in path_fixture_names():
        cfg, seed, cases = load_paths(name)
        for case in cases:
            yield name, cfg, seed, case


def load_helpers():
    return np.load(os.path.join(GOLDEN_DIR, "helpers.npz"), allow_pickle=False)


def load_search():
    with open(os.path.join(GOLDEN_DIR, "search.json")) as f:
        return json.load(f)
