"""Helpers to load tests/golden/*.npz together with the workload that produced them."""
import json
import os

import numpy as np

from oracle.workloads import make_workload

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        return sorted(json.load(f).keys())


def load_case(name):
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        kwargs = json.load(f)[name]
    cfg = make_workload(**kwargs)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    if cfg["N"] == 1 and kwargs.get("name") is None:
        cfg["x"][:] = 0.0
        cfg["y"][:] = 0.0
    # the golden file also stores the inputs: they must be bit-identical to the regenerated ones
    for k in ("x", "y", "actions", "mu0"):
        assert np.array_equal(gold[k], cfg[k]), "workload generator drifted for %s/%s" % (name, k)
    return cfg, gold
