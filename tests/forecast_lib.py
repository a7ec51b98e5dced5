"""Forecast producer test helpers: one scripted measurement sequence driven through the oracle
(oracle/forecast_oracle.hpp), the reference's own forecast.cpp/kalman.cpp build (oracle/_ref/libforecast_ref.so,
only where /root/reference was mounted at build time) or the CUDA producer (mppi_b200_forecast_*)."""
import ctypes as C
import os

import numpy as np

from assistedmanipulation_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libforecast_ref.so")
_dp = C.POINTER(C.c_double)

# (name, type, horison_or_window, time_step, order)
CASES = [
    ("locf", abi.FORECAST_LOCF, 0.4, 0.0, 0),
    ("average", abi.FORECAST_AVERAGE, 0.3, 0.0, 0),
    ("kalman0", abi.FORECAST_KALMAN, 1.0, 0.01, 0),
    ("kalman1", abi.FORECAST_KALMAN, 1.0, 0.01, 1),     # the simulation default (forecast.hpp:397-413)
    ("kalman2", abi.FORECAST_KALMAN, 0.5, 0.02, 2),
]


def script(seed, steps=40):
    """A mixed sequence of update(measurement, time) / update(time) calls and, after each, query times.
    Returns a list of (kind, time, measurement-or-None, queries)."""
    rng = np.random.default_rng(seed)
    out, t = [], 0.0
    for i in range(steps):
        t += float(rng.uniform(0.004, 0.03))
        kind = "time" if rng.uniform() < 0.25 else "measure"
        m = None
        if kind == "measure":
            m = np.concatenate([10.0 * np.sin(0.7 * t + np.arange(3)), rng.normal(0, 2.0, 3)]) + rng.normal(0, 0.05, 6)
        queries = t + np.array([0.0, 0.0037, 0.05, 0.2, 0.39, 0.6, 1.5])
        out.append((kind, t, m, queries))
    return out


class CForecast:
    """oracle_forecast_* or ref_forecast_* (same signatures)."""

    def __init__(self, lib, prefix, typ, hw, dt, order, initial=None):
        self.lib, self.p = lib, prefix
        g = lambda n: getattr(lib, prefix + n)
        g("create").argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint, _dp]
        g("create").restype = C.c_void_p
        g("destroy").argtypes = [C.c_void_p]
        g("update").argtypes = [C.c_void_p, _dp, C.c_int, C.c_double]
        g("update_time").argtypes = [C.c_void_p, C.c_double]
        g("get").argtypes = [C.c_void_p, C.c_double, _dp, C.c_int]
        init = None if initial is None else np.ascontiguousarray(initial, dtype=np.float64)
        self.h = g("create")(typ, 6, hw, dt, order, None if init is None else init.ctypes.data_as(_dp))

    def update(self, m, t):
        m = np.ascontiguousarray(m, dtype=np.float64)
        getattr(self.lib, self.p + "update")(self.h, m.ctypes.data_as(_dp), 6, t)

    def update_time(self, t):
        getattr(self.lib, self.p + "update_time")(self.h, t)

    def forecast(self, t):
        out = np.zeros(6)
        getattr(self.lib, self.p + "get")(self.h, t, out.ctypes.data_as(_dp), 6)
        return out

    def close(self):
        getattr(self.lib, self.p + "destroy")(self.h)


def run_script(fc, steps):
    """-> array [len(steps), n_queries, 6]"""
    out = []
    for kind, t, m, queries in steps:
        if kind == "measure":
            fc.update(m, t)
        else:
            fc.update_time(t)
        out.append([fc.forecast(float(q)) for q in queries])
    return np.array(out)


def ref_available():
    return os.path.exists(REF_PATH)


from assistedmanipulation_b200.forecast import DeviceForecast  # noqa: E402,F401  (the product's wrapper; kept importable from here for the tests)
