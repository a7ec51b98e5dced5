"""SURVEY §8f-2, GPU: DynamicsForecast::forecast on the device (mppi_b200_dynamics_forecast_*) against the oracle
restatement (oracle/dynamics_forecast_oracle.hpp) — every recorded quantity over the whole horizon."""
import ctypes as C

import numpy as np
import pytest

import forecast_lib as fl
import oracle_lib
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu
_dp = C.POINTER(C.c_double)


def make_state(rng, moving=True):
    x = abi.huddled_state(10.0)
    x[:10] += rng.normal(0, 0.05, 10)
    if moving:
        x[12:22] = rng.normal(0, 0.3, 10)
    return x


def oracle_forecaster(olib, dt, horison, wrench, apply):
    olib.oracle_dynamics_forecast_create.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int]
    olib.oracle_dynamics_forecast_create.restype = C.c_void_p
    olib.oracle_dynamics_forecast_destroy.argtypes = [C.c_void_p]
    olib.oracle_dynamics_forecast_steps.argtypes = [C.c_void_p]
    olib.oracle_dynamics_forecast_run.argtypes = [C.c_void_p, _dp, C.c_double]
    olib.oracle_dynamics_forecast_read.argtypes = [C.c_void_p, _dp]
    return olib.oracle_dynamics_forecast_create(dt, horison, wrench, apply)


def compare(got, want):
    """Positions / energies to 1e-9 relative; velocities and accelerations are sums of cancelling terms
    (zero control: a = M^-1 (nle - nle)), compared against the scale of the quantities that produce them."""
    R = abi
    for sl, scale in [(R.DF_JOINT_POSITION, 1.0), (R.DF_POSITION, 1.0), (R.DF_ORIENTATION, 1.0), (R.DF_LINEAR_VELOCITY, 1.0),
                      (R.DF_ANGULAR_VELOCITY, 1.0), (R.DF_WRENCH, 1.0), (R.DF_JACOBIAN, 1.0)]:
        np.testing.assert_allclose(got[..., sl], want[..., sl], rtol=1e-9, atol=1e-9 * scale)
    # accelerations: M^-1 times a difference of ~1e2-sized torques
    np.testing.assert_allclose(got[..., R.DF_LINEAR_ACCELERATION], want[..., R.DF_LINEAR_ACCELERATION], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(got[..., R.DF_ANGULAR_ACCELERATION], want[..., R.DF_ANGULAR_ACCELERATION], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(got[..., R.DF_ENERGY], want[..., R.DF_ENERGY], rtol=1e-9, atol=1e-9)
    assert np.all(got[..., R.DF_JOINT_POWER] == 0.0) and np.all(got[..., R.DF_EXTERNAL_POWER] == 0.0)   # pinocchio_dynamics.hpp:211-223


@pytest.mark.parametrize("mode", ["no_wrench", "kalman", "kalman_applied"])
def test_dynamics_forecast_matches_oracle(mode):
    lib = abi.load_library()
    olib = oracle_lib.load()
    B, dt, horison = 3, 0.01, 0.5
    rng = np.random.default_rng(17)
    apply = int(mode == "kalman_applied")
    dev_w = orc_w = None
    if mode != "no_wrench":
        dev_w = fl.DeviceForecast(abi.FORECAST_KALMAN, 1.0, 0.01, 1, batch=B)
        orc_w = [fl.CForecast(olib, "oracle_forecast_", abi.FORECAST_KALMAN, 1.0, 0.01, 1) for _ in range(B)]
    cfg = abi.DynamicsForecastConfig(batch=B, device=0, time_step=dt, horison=horison, apply_wrench=apply)
    h = C.c_void_p()
    assert lib.mppi_b200_dynamics_forecast_create(C.byref(cfg), dev_w.h if dev_w else None, C.byref(h)) == 0, lib.mppi_b200_dynamics_forecast_last_error(None)
    steps = lib.mppi_b200_dynamics_forecast_steps(h)
    assert steps == 50
    orcs = [oracle_forecaster(olib, dt, horison, orc_w[b].h if orc_w else None, apply) for b in range(B)]
    assert olib.oracle_dynamics_forecast_steps(orcs[0]) == steps
    t = 0.0
    for call in range(3):   # the torque left by the previous forecast enters the next set_state (acceleration at step 0)
        t += 0.05
        if dev_w:
            for k in range(4):
                m = rng.normal(0, 2.0 if apply else 15.0, (B, 6))
                tm = t - 0.04 + 0.01 * k
                assert lib.mppi_b200_forecast_update(dev_w.h, m.ctypes.data_as(_dp), tm) == 0
                for b in range(B):
                    orc_w[b].update(m[b], tm)
        states = np.ascontiguousarray(np.stack([make_state(rng, moving=call > 0) for _ in range(B)]))
        assert lib.mppi_b200_dynamics_forecast_run(h, states.ctypes.data_as(_dp), t) == 0, lib.mppi_b200_dynamics_forecast_last_error(h)
        got = np.zeros((B, steps, abi.DYNAMICS_FORECAST_RECORD))
        assert lib.mppi_b200_dynamics_forecast_read(h, got.ctypes.data_as(_dp), got.nbytes) == 0
        want = np.zeros_like(got)
        for b in range(B):
            olib.oracle_dynamics_forecast_run(orcs[b], states[b].ctypes.data_as(_dp), t)
            olib.oracle_dynamics_forecast_read(orcs[b], want[b].ctypes.data_as(_dp))
        if apply:
            # undamped arm under a constant wrench: the trajectories run away exponentially (|q| reaches 1e2 within
            # 0.5 s) and so does any rounding difference; the first 15 steps carry the comparison
            got, want = got[:, :15], want[:, :15]
        compare(got, want)
        assert np.abs(np.linalg.norm(got[..., abi.DF_ORIENTATION], axis=-1) - 1).max() < 1e-12
        if mode != "no_wrench":
            assert np.abs(got[..., abi.DF_WRENCH]).max() > 0.5
        if mode == "kalman_applied" and call == 0:   # the wrench really moves the (initially resting) arm
            assert np.abs(got[:, -1, abi.DF_JOINT_POSITION] - states[:, :12]).max() > 1e-4
    for o in orcs:
        olib.oracle_dynamics_forecast_destroy(o)
    lib.mppi_b200_dynamics_forecast_destroy(h)
    if dev_w:
        dev_w.close()
        for o in orc_w:
            o.close()


def test_dynamics_forecast_create_errors():
    lib = abi.load_library()
    h = C.c_void_p()
    cfg = abi.DynamicsForecastConfig(batch=1, device=0, time_step=0.01, horison=0.0, apply_wrench=0)
    assert lib.mppi_b200_dynamics_forecast_create(C.byref(cfg), None, C.byref(h)) == abi.ERR_INVALID
    assert b"time horison is too small for time step" in lib.mppi_b200_dynamics_forecast_last_error(None)   # dynamics.cpp:71-75
    w = fl.DeviceForecast(abi.FORECAST_LOCF, 1.0, 0.0, 0, batch=2)
    cfg = abi.DynamicsForecastConfig(batch=3, device=0, time_step=0.01, horison=0.1, apply_wrench=0)
    assert lib.mppi_b200_dynamics_forecast_create(C.byref(cfg), w.h, C.byref(h)) == abi.ERR_INVALID
    w.close()


def test_facade_dynamics_forecast_class(tmp_path):
    """tests/cpp/forecast_demo.cpp drives FrankaRidgeback::DynamicsForecast (facade, reference API) like the Actor;
    the printed trajectory points must be the oracle's."""
    import subprocess
    from test_abi_cpu import build_facade_demo
    exe = build_facade_demo("forecast_demo")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = {line.split()[0]: [float(v) for v in line.split()[1:]] for line in r.stdout.splitlines() if line.startswith("d")}
    olib = oracle_lib.load()
    w = fl.CForecast(olib, "oracle_forecast_", abi.FORECAST_LOCF, 1e9, 0.0, 0, initial=np.zeros(6))
    w.update(np.array([3.0, -1.0, 2.0, 0.1, 0.2, -0.3]), 0.02)
    w.update_time(0.03)
    o = oracle_forecaster(olib, 0.01, 0.3, w.h, 0)
    x = abi.huddled_state(10.0)
    x[12 + 4] = 0.3
    olib.oracle_dynamics_forecast_run(o, x.ctypes.data_as(_dp), 0.05)
    want = np.zeros((30, abi.DYNAMICS_FORECAST_RECORD))
    olib.oracle_dynamics_forecast_read(o, want.ctypes.data_as(_dp))
    assert rows["dynforecast"] == [30.0, 0.05, 0.01]
    np.testing.assert_allclose(rows["df_first"], want[0, abi.DF_POSITION], rtol=1e-9)
    np.testing.assert_allclose(rows["df_last"], want[29, abi.DF_POSITION], rtol=1e-9)
    np.testing.assert_allclose(rows["df_wrench"], [3.0, -1.0, 2.0])
    np.testing.assert_allclose(rows["df_energy"], [want[0, abi.DF_ENERGY], want[29, abi.DF_ENERGY]], rtol=1e-9)
    olib.oracle_dynamics_forecast_parameterise.argtypes = [C.c_void_p, C.c_double]
    olib.oracle_dynamics_forecast_parameterise.restype = C.c_long
    assert olib.oracle_dynamics_forecast_parameterise(o, 0.0) == 0 and olib.oracle_dynamics_forecast_parameterise(o, 0.175) == 12
    np.testing.assert_allclose(rows["df_lookup"], [want[0, 12], want[12, 12]], rtol=1e-9)
    np.testing.assert_allclose(rows["df_q4"], [want[0, 4], want[29, 4]], rtol=1e-9)
    assert rows["df_q4"][1] != rows["df_q4"][0]   # the moving joint coasts
    olib.oracle_dynamics_forecast_destroy(o)
    w.close()
