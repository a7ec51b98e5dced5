"""SURVEY §8f-1, CPU: the forecast oracle (oracle/forecast_oracle.hpp) against golden vectors produced by the
reference's own forecast.cpp / kalman.cpp (tests/golden/ref_forecast.npz, tools/gen_forecast_golden.py) and,
where it is built, against that reference build live. Bit exact: same IEEE operations in the same order."""
import ctypes as C

import numpy as np
import pytest

import forecast_lib as fl
import oracle_lib


@pytest.fixture(scope="module")
def golden():
    return np.load(fl.os.path.join(fl.ROOT, "tests", "golden", "ref_forecast.npz"))


@pytest.fixture(scope="module")
def olib():
    return oracle_lib.load()


@pytest.mark.parametrize("case", fl.CASES, ids=[c[0] for c in fl.CASES])
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_matches_reference_golden(olib, golden, case, seed):
    name, typ, hw, dt, order = case
    fc = fl.CForecast(olib, "oracle_forecast_", typ, hw, dt, order)
    got = fl.run_script(fc, fl.script(seed))
    fc.close()
    np.testing.assert_array_equal(got, golden["%s_s%d" % (name, seed)])


def average_known_answers(fc):
    """The reference's own AverageForecast test sequence, src/test/case/forecast.cpp:62-100."""
    pad = lambda v: np.array(list(v) + [0, 0, 0], dtype=float)
    rows = [fc.forecast(0.0)]
    fc.update(pad((0, 1.0, 0)), 1.01); rows.append(fc.forecast(5.0))
    fc.update(pad((0, 1.5, 0)), 1.5); rows.append(fc.forecast(10.0))
    fc.update(pad((1.0, 1.0, 1.0)), 3.0); rows.append(fc.forecast(3.0))
    for i in range(10):
        fc.update(pad((i, i, i)), 4.5 + i * 0.05)
    rows.append(fc.forecast(3.5))
    fc.update_time(10.0); rows.append(fc.forecast(10.0))
    return np.array(rows)


def test_average_known_answers(olib, golden):
    fc = fl.CForecast(olib, "oracle_forecast_", 1, 1.0, 0.0, 0)
    rows = average_known_answers(fc)
    fc.close()
    # the expectations written in the reference test (the last one, (9,9,9), is not what its code returns)
    expect = [(0, 0, 0), (0, 1, 0), (0, 1.25, 0), (1, 1, 1), (4.5, 4.5, 4.5)]
    np.testing.assert_allclose(rows[:5, :3], expect, rtol=1e-12)
    np.testing.assert_array_equal(rows, golden["average_kat"])


def test_locf_carries_forward(olib):
    """src/test/case/forecast.cpp:23-60 with the default horison: the observation is carried forward."""
    rng = np.random.default_rng(3)
    fc = fl.CForecast(olib, "oracle_forecast_", 0, 1e30, 0.0, 0, initial=[1, 2, 3, 0, 0, 0])
    for _ in range(5):
        m = rng.uniform(-1, 1, 6)
        fc.update(m, 0.0)
        for q in (0.0, 1.0, 2.0):
            np.testing.assert_array_equal(fc.forecast(q), m)
    fc.close()


def test_kalman_tracks_a_line(olib):
    """src/test/case/forecast.cpp:103-160 in spirit: a first-order filter fed an exact line extrapolates it."""
    fc = fl.CForecast(olib, "oracle_forecast_", 2, 3.0, 0.1, 1)
    slope = np.array([1.0, -2.0, 0.5, 0.0, 3.0, -1.0])
    for i in range(60):
        fc.update(slope * (0.1 * i), 0.1 * i)
    t = 0.1 * 59
    for ahead in (0.0, 1.0, 2.0):
        np.testing.assert_allclose(fc.forecast(t + ahead), slope * (t + ahead), rtol=0.05, atol=0.05)
    fc.close()


@pytest.mark.skipif(not fl.ref_available(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("case", fl.CASES, ids=[c[0] for c in fl.CASES])
def test_oracle_matches_reference_build_live(olib, case):
    name, typ, hw, dt, order = case
    ref = C.CDLL(fl.REF_PATH)
    steps = fl.script(11, steps=80)
    a = fl.CForecast(olib, "oracle_forecast_", typ, hw, dt, order)
    b = fl.CForecast(ref, "ref_forecast_", typ, hw, dt, order)
    np.testing.assert_array_equal(fl.run_script(a, steps), fl.run_script(b, steps))
    a.close(); b.close()
