"""Edge cases of the path against the oracle: degenerate sizes, ragged horizons, extreme time shifts,
discounting, unbounded / unsmoothed controls, a full (non-diagonal) sampling covariance."""
import numpy as np
import pytest

import oracle_lib as ol
from assistedmanipulation_b200 import abi
from test_gpu_parity import run_pair

pytestmark = pytest.mark.gpu

TOY = (abi.SYSTEM_TOY, abi.OBJECTIVE_TOY)


def test_single_configured_rollout(oracle):
    # rollouts = 1 -> three rollouts in total (zero noise, -U_prev, one sample): mppi.hpp:306
    run_pair(oracle, *TOY, abi.default_toy_objective(), 1, 0.2, np.array([0.3, 0.1, 0.0, 0.0]), 4, 0.05)


def test_single_step_horizon(oracle):
    # T = 1: horison == time_step; the smoothing window is 2w+2 long and every update shifts the whole horizon out
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 62, 0.01, abi.huddled_state(), 3, 0.01, keep=4)


def test_ragged_horizon_is_rounded_up(oracle):
    # T = ceil(horison / time_step) in double (mppi.cpp:85): 0.305 / 0.01 -> 31 steps
    import engine_lib as el
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 8, 0.305)
    e = el.Engine(h, abi.default_toy_objective())
    o = ol.Oracle(oracle, h, abi.default_toy_objective())
    assert e.query(abi.QUERY_STEP_COUNT) == o.query(abi.QUERY_STEP_COUNT) == 31
    e.close()
    o.close()
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 30, 0.315, abi.huddled_state(), 3, 0.05, keep=5)


def test_keep_every_rollout(oracle):
    # keep_best == rollouts: nothing is resampled in full, only the tails the shift exposes
    run_pair(oracle, *TOY, abi.default_toy_objective(), 40, 0.3, np.zeros(4), 5, 0.05, keep=40)


def test_shift_by_the_whole_horizon(oracle):
    # shift_by == T: every column becomes the last column of the previous optimum (mppi.cpp:204-206), kept rows are all tail
    run_pair(oracle, *TOY, abi.default_toy_objective(), 50, 0.2, np.zeros(4), 3, 0.2, keep=10)


def test_updates_without_time_advance(oracle):
    # shift_by == 0 ("subsample update", mppi.cpp:193): kept rollouts are left untouched, nothing shifts
    run_pair(oracle, *TOY, abi.default_toy_objective(), 50, 0.3, np.array([0.0, 0.5, 0.0, 0.0]), 4, 0.0, keep=10)
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 40, 0.2, abi.huddled_state(), 3, 0.004, keep=10)


def test_discounted_cost(oracle):
    # std::pow(cost_discount_factor, step) (mppi.cpp:326)
    import engine_lib as el
    K, T = 60, 30
    for system, objective, params, x0, nu in ((abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective(), np.array([0.2, 0.0, 0.0, 0.1]), 2),
                                               (abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), abi.huddled_state(), 12)):
        h = abi.make_config(system, objective, K, 0.3, discount=0.93, dynamics_mode=abi.DYNAMICS_FUSED)
        o, e = ol.Oracle(oracle, h, params), el.Engine(h, params)
        eps = np.random.default_rng(1).standard_normal((K + 2, T, nu)) * (1.0 if nu == 2 else np.sqrt(abi.FRANKA_COVARIANCE_DIAG))
        for u in range(2):
            assert o.update(x0, 0.05 * u, None, eps) == 0 and e.update(x0, 0.05 * u, None, eps) == 0
            co, ce = o.read(abi.READ_COSTS, K + 2), e.read(abi.READ_COSTS, K + 2)
            assert (np.abs(ce - co) / np.abs(co)).max() <= 1e-9
            Uo = o.read(abi.READ_OPTIMAL, nu * T)
            assert np.abs(e.read(abi.READ_OPTIMAL, nu * T) - Uo).max() <= 1e-9 * np.abs(Uo).max()
        o.close()
        e.close()


def test_unbounded_unsmoothed_no_default(oracle):
    # control_bound = false, smoothing = nullopt, control_default = nullopt: get() past the horizon returns the last column (mppi.cpp:496-501)
    import engine_lib as el
    K, T = 30, 10
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.1, control_bound=False, smoothing=None, gradient_step=50.0)
    o, e = ol.Oracle(oracle, h, abi.default_toy_objective()), el.Engine(h, abi.default_toy_objective())
    eps = np.random.default_rng(2).standard_normal((K + 2, T, 2)) * 3
    assert o.update(np.zeros(4), 0.0, None, eps) == 0 and e.update(np.zeros(4), 0.0, None, eps) == 0
    Uo, Ue = o.read(abi.READ_OPTIMAL, 2 * T), e.read(abi.READ_OPTIMAL, 2 * T)
    assert np.abs(Uo).max() > 5.0 and np.abs(Ue - Uo).max() <= 1e-9 * np.abs(Uo).max()   # beyond the +-5 bounds: not clamped
    assert np.allclose(e.get(5.0), Uo.reshape(T, 2)[-1], rtol=1e-9) and np.allclose(o.get(5.0), Uo.reshape(T, 2)[-1])
    hd = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.1, control_default=np.array([0.25, -0.5]))
    ed = el.Engine(hd, abi.default_toy_objective())
    assert ed.update(np.zeros(4), 0.0, None, eps) == 0
    assert np.array_equal(ed.get(5.0), [0.25, -0.5])
    for x in (o, e, ed):
        x.close()


def test_full_covariance_philox():
    # eps = V sqrt(Lambda) z (gaussian.hpp:48-55,70-75): sample covariance of the generated noise ~ the configured one
    import engine_lib as el
    K, T = 20000, 8
    cov = np.array([[2.0, 0.6], [0.6, 0.5]])
    e = el.Engine(abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.08, covariance=cov), abi.default_toy_objective())
    assert e.update(np.zeros(4), 0.0, seed=3) == 0
    z = e.read(abi.READ_NOISE, (K + 2) * T * 2).reshape(K + 2, T, 2)[2:].reshape(-1, 2)
    assert np.abs(np.cov(z.T) - cov).max() < 0.02 and np.abs(z.mean(axis=0)).max() < 0.01
    e.close()


def test_full_covariance_oracle_transform_matches(oracle):
    # the engine's V sqrt(Lambda) equals the oracle's (same eigen ordering): inject z through both with identity noise
    import engine_lib as el
    K, T = 16, 6
    cov = np.array([[2.0, 0.6], [0.6, 0.5]])
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.06, covariance=cov, smoothing=None)
    o, e = ol.Oracle(oracle, h, abi.default_toy_objective()), el.Engine(h, abi.default_toy_objective())
    eps = np.random.default_rng(6).standard_normal((K + 2, T, 2))
    assert o.update(np.zeros(4), 0.0, None, eps) == 0 and e.update(np.zeros(4), 0.0, None, eps) == 0   # injected noise bypasses the transform on both sides
    Uo = o.read(abi.READ_OPTIMAL, 2 * T)
    assert np.abs(e.read(abi.READ_OPTIMAL, 2 * T) - Uo).max() <= 1e-9 * np.abs(Uo).max()
    o.close()
    e.close()


def test_get_from_another_thread_during_updates():
    """mppi.cpp:178-182,492: the control loop reads the published sequence while the controller updates. A reader thread
    hammers mppi_b200_get during 40 updates: every answer is the interpolation of ONE published sequence (the one before
    or the one after the concurrent update, never a mixture), and the reader never waits for a whole update."""
    import threading
    import time as clock
    import engine_lib as el
    K, T, nu = 16384, 64, 12
    e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.64, dynamics_mode=abi.DYNAMICS_FUSED), abi.default_track_point())
    x0 = abi.huddled_state()
    assert e.update(x0, 0.0, seed=1) == 0
    published = {0.0: e.read(abi.READ_OPTIMAL, nu * T).reshape(T, nu)}
    samples, stop, query_time = [], [False], [0.013]

    def reader():
        out = np.zeros(nu)
        while not stop[0]:
            tq = query_time[0]
            t0 = clock.perf_counter()
            rc = e.lib.mppi_b200_get(e.h, el.ptr(out), tq)
            samples.append((tq, rc, out.copy(), clock.perf_counter() - t0))

    th = threading.Thread(target=reader)
    th.start()
    update_s = []
    for u in range(1, 41):
        t = 0.05 * u
        t0 = clock.perf_counter()
        assert e.update(x0, t, seed=1) == 0, e.error()
        update_s.append(clock.perf_counter() - t0)
        published[t] = np.frombuffer(bytes(e.read(abi.READ_OPTIMAL, nu * T)), dtype=np.float64).reshape(T, nu).copy()
        query_time[0] = t + 0.013
    stop[0] = True
    th.join()
    e.close()
    assert len(samples) > 100
    times = sorted(published)
    checked = 0
    for tq, rc, got, _ in samples:
        if rc != 0:
            continue   # asked for a time before the sequence published meanwhile: refused like the reference's assert
        ok = False
        for tn in times:
            if tn > tq:
                break
            s = (tq - tn) / 0.01
            lo = int(s)
            if lo + 1 >= T:
                continue
            w = s - lo
            ok = ok or np.array_equal(got, (1.0 - w) * published[tn][lo] + w * published[tn][lo + 1])
        assert ok, tq
        checked += 1
    assert checked > 50
    # the reader is held up by the publication copy only, not by an update (the median update here takes ~0.5 ms)
    waits = np.array([s[3] for s in samples])
    assert np.median(waits) < 0.2 * np.median(update_s)


def test_host_buffers_refilled_in_place_and_strided_views():
    """The Python host side keeps the ctypes pointer of a state / wrench array for as long as the caller passes the same
    object (engine.py: _host_pointer). A buffer refilled in place must be read afresh by every update, and arrays that need
    a conversion (strided views, Fortran order) are converted on every call."""
    import engine_lib as el
    import cases
    K, T, nu = 256, 16, 12
    params, W = cases.assisted_params(True, abi.LINKS_BODY_COM), cases.constant_wrench(T)
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 0.16, dynamics_mode=abi.DYNAMICS_FUSED)
    a, b = el.Engine(h, params), el.Engine(h, params)
    state, wrench = abi.huddled_state(10.0), W.copy()                       # reused by `a`
    columns = np.zeros((state.size, 2))
    for u in range(4):
        state[3] = 0.1 * u; wrench[:, 0] = 1.0 + u                           # refilled in place
        assert a.update(state, 0.05 * u, wrench, seed=9) == 0, a.error()
        if u % 2:                                                            # the same values through arrays that need a copy
            columns[:, 0] = state
            assert b.update(columns[:, 0], 0.05 * u, np.asfortranarray(wrench), seed=9) == 0, b.error()
        else:
            assert b.update(state.copy(), 0.05 * u, wrench.copy(), seed=9) == 0, b.error()
        assert np.array_equal(a.read(abi.READ_OPTIMAL, nu * T), b.read(abi.READ_OPTIMAL, nu * T)), u
        assert np.array_equal(a.read(abi.READ_COSTS, K + 2), b.read(abi.READ_COSTS, K + 2)), u
    a.close(); b.close()
