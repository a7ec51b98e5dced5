"""SURVEY §8f-1, GPU: the batched CUDA forecast producer (mppi_b200_forecast_*) against the oracle and the
reference-generated golden vectors — bit exact (the kernels run the reference's IEEE operations in order) —
and the device-resident hand-over of its table to a batched engine."""
import ctypes as C

import numpy as np
import pytest

import forecast_lib as fl
import oracle_lib
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(fl.os.path.join(fl.ROOT, "tests", "golden", "ref_forecast.npz"))


@pytest.mark.parametrize("case", fl.CASES, ids=[c[0] for c in fl.CASES])
@pytest.mark.parametrize("seed", [1, 2])
def test_device_matches_reference_golden(golden, case, seed):
    name, typ, hw, dt, order = case
    fc = fl.DeviceForecast(typ, hw, dt, order)
    got = fl.run_script(fc, fl.script(seed))
    fc.close()
    np.testing.assert_array_equal(got, golden["%s_s%d" % (name, seed)])


@pytest.mark.parametrize("case", fl.CASES, ids=[c[0] for c in fl.CASES])
def test_device_table_matches_oracle_batched(case):
    """Every forecaster of a batch gets its own measurements; the table is the oracle's forecast(time + k dt)."""
    name, typ, hw, dt, order = case
    B, T, step = 5, 30, 0.015
    olib = oracle_lib.load()
    rng = np.random.default_rng(5)
    init = rng.normal(0, 1, (B, 6)) if typ != abi.FORECAST_AVERAGE else None
    dev = fl.DeviceForecast(typ, hw, dt, order, batch=B, initial=init)
    orc = [fl.CForecast(olib, "oracle_forecast_", typ, hw, dt, order, initial=None if init is None else init[b]) for b in range(B)]
    t = 0.0
    for i in range(25):
        t += float(rng.uniform(0.005, 0.02))
        if i % 6 == 5:
            dev.update_time(t)
            for o in orc:
                o.update_time(t)
        else:
            m = rng.normal(0, 5, (B, 6))
            assert dev.lib.mppi_b200_forecast_update(dev.h, m.ctypes.data_as(fl._dp), t) == 0
            for b, o in enumerate(orc):
                o.update(m[b], t)
        table = dev.table(t, step, T)
        want = np.array([[o.forecast(t + k * step) for k in range(T)] for o in orc])
        np.testing.assert_array_equal(table, want)
    dev.close()
    for o in orc:
        o.close()


def test_average_window_overflow_is_reported():
    fc = fl.DeviceForecast(abi.FORECAST_AVERAGE, 1e9, 0.0, 0)
    for i in range(2100):
        fc.update(np.ones(6), 0.001 * i)
    out = np.zeros((1, 1, 6))
    rc = fc.lib.mppi_b200_forecast_table(fc.h, 3.0, 1.0, 1, out.ctypes.data_as(fl._dp))
    assert rc == abi.ERR_UNSUPPORTED
    assert b"2048" in fc.lib.mppi_b200_forecast_last_error(fc.h)
    fc.close()


def test_create_errors():
    lib = abi.load_library()
    h = C.c_void_p()
    bad = abi.ForecastConfig(type=abi.FORECAST_AVERAGE, batch=1, device=0, order=0, time_step=0.0, horison=0.0, window=-1.0)
    assert lib.mppi_b200_forecast_create(C.byref(bad), None, C.byref(h)) == abi.ERR_INVALID
    assert b"negative" in lib.mppi_b200_forecast_last_error(None)   # forecast.cpp:44-47
    bad = abi.ForecastConfig(type=abi.FORECAST_KALMAN, batch=1, device=0, order=3, time_step=0.01, horison=1.0, window=0.0)
    assert lib.mppi_b200_forecast_create(C.byref(bad), None, C.byref(h)) == abi.ERR_INVALID


def test_engine_consumes_device_table():
    """A batched engine fed the producer's device table computes what it computes from the same table on the host."""
    from cases import assisted_params
    from engine_lib import Engine
    B, K, T = 3, 64, 20
    rng = np.random.default_rng(9)
    fc = fl.DeviceForecast(abi.FORECAST_KALMAN, 1.0, 0.01, 1, batch=B)
    t = 0.0
    for i in range(10):
        t += 0.01
        m = rng.normal(0, 20, (B, 6))
        assert fc.lib.mppi_b200_forecast_update(fc.h, m.ctypes.data_as(fl._dp), t) == 0
    holder = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, T * 0.01, keep_best=20,
                             dynamics_mode=abi.DYNAMICS_FUSED, batch=B)
    objective = assisted_params()
    x = np.tile(abi.huddled_state(), (B, 1))
    host_table = fc.table(t, 0.01, T)
    results = []
    for mode in ("host", "device"):
        eng = Engine(holder, objective)
        if mode == "device":
            assert eng.lib.mppi_b200_set_wrench_device(eng.h, fc.table_device(t, 0.01, T)) == 0
        for u in range(3):
            assert eng.update(x, 0.0, wrench=host_table if mode == "host" else None, seed=7) == 0, eng.error()
        results.append((eng.read(abi.READ_COSTS, B * (K + 2)), eng.read(abi.READ_OPTIMAL, B * T * 12)))
        eng.close()
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_array_equal(results[0][1], results[1][1])
    assert np.std(host_table) > 0
    fc.close()


def test_facade_forecast_classes_follow_the_reference_tests(golden):
    """tests/cpp/forecast_demo.cpp: the reference's forecast test sequences through the facade classes."""
    import subprocess
    from test_abi_cpu import build_facade_demo
    exe = build_facade_demo("forecast_demo")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "prediction window time is negative" in r.stderr and "kalman forecast selected with no configuration provided" in r.stderr
    rows = {}
    for line in r.stdout.splitlines():
        tag, *vals = line.split()
        rows.setdefault(tag, []).append([float(v) for v in vals])
    locf = np.array(rows["locf"]).reshape(3, 3, 3)
    samples = np.array([[0.1, -0.7, 0.3], [0.9, 0.2, -0.4], [-0.5, 0.6, 0.8]])
    for i in range(3):
        np.testing.assert_array_equal(locf[i], np.tile(samples[i], (3, 1)))   # carried forward
    np.testing.assert_array_equal(np.array(rows["average"]), golden["average_kat"][:, :3])
    # the same Kalman sequence through the oracle
    olib = oracle_lib.load()
    o = fl.CForecast(olib, "oracle_forecast_", 2, 3.0, 0.1, 1)
    slope = np.array([1.0, -2.0, 0.5, 0.0, 3.0, -1.0])
    for i in range(60):
        t = 0.1 * i
        o.update(slope * t, t)
    want = [o.forecast(t), o.forecast(t + 1.0), o.forecast(t + 2.0)]
    o.update_time(t + 0.05)
    want.append(o.forecast(t + 0.5))
    np.testing.assert_array_equal(np.array(rows["kalman"]), np.array(want))
    np.testing.assert_allclose(np.array(rows["kalman"])[1], slope * (t + 1.0), rtol=0.05, atol=0.05)
    o.close()
