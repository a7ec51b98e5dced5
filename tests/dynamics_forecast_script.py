"""One scripted DynamicsForecast session driven through the REFERENCE's own frankaridgeback/dynamics.cpp
(oracle/_ref/libmppi_ref.so: ref_dynamics_forecast_*) or through the oracle restatement
(oracle/dynamics_forecast_oracle.hpp): wrench observations -> forecast(state, time) -> the recorded horizon, three
consecutive calls (the torque left by one forecast enters the next set_state), plus parameterise() probes."""
import ctypes as C

import numpy as np

from assistedmanipulation_b200 import abi

_dp = C.POINTER(C.c_double)
DT, HORISON = 0.01, 0.2
# (name, forecaster type, horison / window, forecaster time step, order)
CASES = [("kalman1", abi.FORECAST_KALMAN, 1.0, 0.01, 1), ("locf", abi.FORECAST_LOCF, 0.4, 0.0, 0), ("average", abi.FORECAST_AVERAGE, 0.3, 0.0, 0)]
PROBE_TIMES = [-1.0, 0.0, 0.05, 0.1, 0.149, 0.15, 0.1501, 0.163, 0.175, 0.19999, 0.2, 0.3, 5.0]


def session(seed):
    rng = np.random.default_rng(seed)
    calls, t = [], 0.0
    for call in range(3):
        t += 0.05
        obs = [(t - 0.04 + 0.01 * k, rng.normal(0, 15.0, 6)) for k in range(4)]
        x = abi.huddled_state(10.0)
        x[:10] += rng.normal(0, 0.05, 10)
        if call > 0:
            x[12:22] = rng.normal(0, 0.3, 10)
        calls.append((t, obs, x))
    return calls


class Reference:
    def __init__(self, ref, typ, hw, fdt, order):
        self.ref = ref
        ref.ref_dynamics_forecast_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_uint]
        ref.ref_dynamics_forecast_create.restype = C.c_void_p
        for n in ("destroy", "steps", "observe_time"):
            getattr(ref, "ref_dynamics_forecast_" + n).argtypes = [C.c_void_p] + ([C.c_double] if n == "observe_time" else [])
        ref.ref_dynamics_forecast_observe.argtypes = [C.c_void_p, _dp, C.c_double]
        ref.ref_dynamics_forecast_run.argtypes = [C.c_void_p, _dp, C.c_double]
        ref.ref_dynamics_forecast_read.argtypes = [C.c_void_p, _dp]
        ref.ref_dynamics_forecast_parameterise.argtypes = [C.c_void_p, C.c_double]
        ref.ref_dynamics_forecast_parameterise.restype = C.c_long
        self.h = ref.ref_dynamics_forecast_create(DT, HORISON, typ, hw, fdt, order)
        self.steps = ref.ref_dynamics_forecast_steps(self.h)

    def observe(self, m, t):
        m = np.ascontiguousarray(m)
        self.ref.ref_dynamics_forecast_observe(self.h, m.ctypes.data_as(_dp), t)

    def run(self, x, t):
        x = np.ascontiguousarray(x)
        self.ref.ref_dynamics_forecast_run(self.h, x.ctypes.data_as(_dp), t)
        out = np.zeros((self.steps, abi.DYNAMICS_FORECAST_RECORD))
        self.ref.ref_dynamics_forecast_read(self.h, out.ctypes.data_as(_dp))
        return out

    def parameterise(self, t):
        return self.ref.ref_dynamics_forecast_parameterise(self.h, t)

    def close(self):
        self.ref.ref_dynamics_forecast_destroy(self.h)


class Oracle:
    def __init__(self, olib, typ, hw, fdt, order):
        import forecast_lib as fl
        self.olib = olib
        self.w = fl.CForecast(olib, "oracle_forecast_", typ, hw, fdt, order)
        olib.oracle_dynamics_forecast_create.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int]
        olib.oracle_dynamics_forecast_create.restype = C.c_void_p
        olib.oracle_dynamics_forecast_destroy.argtypes = [C.c_void_p]
        olib.oracle_dynamics_forecast_steps.argtypes = [C.c_void_p]
        olib.oracle_dynamics_forecast_run.argtypes = [C.c_void_p, _dp, C.c_double]
        olib.oracle_dynamics_forecast_read.argtypes = [C.c_void_p, _dp]
        olib.oracle_dynamics_forecast_parameterise.argtypes = [C.c_void_p, C.c_double]
        olib.oracle_dynamics_forecast_parameterise.restype = C.c_long
        self.h = olib.oracle_dynamics_forecast_create(DT, HORISON, self.w.h, 0)
        self.steps = olib.oracle_dynamics_forecast_steps(self.h)

    def observe(self, m, t):
        self.w.update(m, t)

    def run(self, x, t):
        x = np.ascontiguousarray(x)
        self.olib.oracle_dynamics_forecast_run(self.h, x.ctypes.data_as(_dp), t)
        out = np.zeros((self.steps, abi.DYNAMICS_FORECAST_RECORD))
        self.olib.oracle_dynamics_forecast_read(self.h, out.ctypes.data_as(_dp))
        return out

    def parameterise(self, t):
        return self.olib.oracle_dynamics_forecast_parameterise(self.h, t)

    def close(self):
        self.olib.oracle_dynamics_forecast_destroy(self.h)
        self.w.close()


def run_session(f, seed):
    records, index = [], []
    for t, obs, x in session(seed):
        for tm, m in obs:
            f.observe(m, tm)
        records.append(f.run(x, t))
        index.append([f.parameterise(q) for q in PROBE_TIMES])
    return np.stack(records), np.array(index)
