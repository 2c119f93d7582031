"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from dqmc_oracle import SdwParams, HubbardParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_DERIVED = ("N", "beta")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def sdw_params_of(g):
    d = json.loads(str(g["params"]))
    for k in _DERIVED:
        d.pop(k, None)
    return SdwParams(**d)


def hubbard_params_of(g):
    d = json.loads(str(g["params"]))
    d.pop("N", None)
    return HubbardParams(**d)


def maxabs(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
