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


def big_case_names():
    """Reference-generated vectors at the BASELINE.json training-set sizes (oracle/make_golden.py --large): noise 1e-5 at
    N = 200 / 500, i.e. cond(K + noise I) ~ 1e6 .. 1e7.  The (E, N, N) inverse is not stored (diagonal + row sums are)."""
    with open(os.path.join(GOLDEN_DIR, "cases_big.json")) as f:
        return sorted(json.load(f).keys())


def load_big_case(name):
    with open(os.path.join(GOLDEN_DIR, "cases_big.json")) as f:
        kwargs = json.load(f)[name]
    cfg = make_workload(**kwargs)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, "big_" + name + ".npz")))
    for k in ("x", "y", "actions", "mu0"):
        assert np.array_equal(gold[k], cfg[k]), "workload generator drifted for %s/%s" % (name, k)
    return cfg, gold
