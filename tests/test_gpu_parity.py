"""Parity of the CUDA engine (through the C ABI) against the CPU oracle on identical injected
noise. Tolerances are BASELINE.json's: argmin and kept indices bit-exact; per-rollout cost
<= 1e-9 relative in FP64; updated control sequence <= 1e-9 relative (FP64) / <= 1e-4 (FP32 fast mode).
"""
import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu

COST_RTOL_FP64 = 1e-9
U_RTOL_FP64 = 1e-9
U_RTOL_FP32 = 1e-4


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def run_pair(oracle, system, objective, params, K, horison, x0, updates, cadence, *, keep=0, wrench=None, precision=abi.FP64,
             mode=abi.DYNAMICS_FUSED, smoothing=(10, 1), seed=11, cost_rtol=COST_RTOL_FP64, u_rtol=U_RTOL_FP64, threads=8):
    import engine_lib as el
    holder_o = abi.make_config(system, objective, K, horison, keep_best=keep, smoothing=smoothing, threads=threads)
    holder_e = abi.make_config(system, objective, K, horison, keep_best=keep, smoothing=smoothing, precision=precision, dynamics_mode=mode)
    o = ol.Oracle(oracle, holder_o, params)
    e = el.Engine(holder_e, params)
    try:
        nu, T, R = holder_o.cfg.control_dof, o.query(abi.QUERY_STEP_COUNT), o.query(abi.QUERY_ROLLOUT_COUNT)
        assert (e.query(abi.QUERY_STEP_COUNT), e.query(abi.QUERY_ROLLOUT_COUNT)) == (T, R)
        rng = np.random.default_rng(seed)
        sigma = np.sqrt(np.diag(np.array([[holder_o.cfg.covariance[c * nu + r] for c in range(nu)] for r in range(nu)])))
        for u in range(updates):
            t = u * cadence
            eps = rng.standard_normal((R, T, nu)) * sigma
            if precision == abi.FP32:
                eps = eps.astype(np.float32).astype(np.float64)  # both sides see exactly representable noise
            assert o.update(x0, t, wrench, eps) == 0
            assert e.update(x0, t, wrench, eps) == 0, e.error()
            assert e.query(abi.QUERY_SHIFT_BY) == o.query(abi.QUERY_SHIFT_BY)
            co, ce = o.read(abi.READ_COSTS, R), e.read(abi.READ_COSTS, R)
            assert np.array_equal(np.isnan(co), np.isnan(ce))
            ok = ~np.isnan(co)
            if precision == abi.FP32:
                # at most one rollout on the other side of a barrier step (see cases.fp32_costs); the control sequence
                # below must meet the tolerance regardless
                cases.fp32_costs(ce, co, cost_rtol, max_flips=1)
            else:
                assert (np.abs(ce[ok] - co[ok]) / np.abs(co[ok])).max() <= cost_rtol, (u, (np.abs(ce[ok] - co[ok]) / np.abs(co[ok])).max())
            if precision == abi.FP64:
                assert e.query(abi.QUERY_ARGMIN) == o.query(abi.QUERY_ARGMIN)
                if keep:
                    assert np.array_equal(e.read(abi.READ_KEPT, keep, np.int64), o.read(abi.READ_KEPT, keep, np.int64))
                ne, no = e.read(abi.READ_NOISE, R * T * nu).reshape(R, -1), o.read(abi.READ_NOISE, R * T * nu).reshape(R, -1)
                assert np.array_equal(ne[0], no[0]) and np.array_equal(ne[2:], no[2:])  # sampled / kept / shifted rows: bit exact
                assert np.allclose(ne[1], no[1], rtol=u_rtol, atol=u_rtol * max(np.abs(no[1]).max(), 1e-300))  # row 1 = -U_prev
                mm_o, mm_e = o.read(abi.READ_MINMAX, 2), e.read(abi.READ_MINMAX, 2)
                assert np.allclose(mm_e, mm_o, rtol=cost_rtol, atol=0)
            Uo, Ue = o.read(abi.READ_OPTIMAL, nu * T), e.read(abi.READ_OPTIMAL, nu * T)
            assert rel(Ue, Uo) <= u_rtol, (u, rel(Ue, Uo))
            assert rel(e.read(abi.READ_WEIGHTS, R), o.read(abi.READ_WEIGHTS, R)) <= max(u_rtol, 1e-9)
            assert rel(e.read(abi.READ_GRADIENT, nu * T), o.read(abi.READ_GRADIENT, nu * T)) <= max(u_rtol, 1e-9)
            oc_o, oc_e = o.read(abi.READ_OPTIMAL_COST, 1)[0], e.read(abi.READ_OPTIMAL_COST, 1)[0]
            assert abs(oc_e - oc_o) <= max(cost_rtol, u_rtol * 10) * abs(oc_o), (oc_e, oc_o)
            assert np.allclose(e.get(t + 0.013), o.get(t + 0.013), rtol=u_rtol, atol=u_rtol * np.abs(Uo).max())
            if precision == abi.FP32:
                # every update of the fast mode is compared FROM IDENTICAL INPUTS: the oracle continues from the sequence the
                # engine published (1e-5 apart after one update, which would otherwise move rollouts across barrier steps
                # through the inputs rather than through the arithmetic under test)
                o.set_optimal(Ue)
    finally:
        o.close()
        e.close()


def test_toy_config1_full_size(oracle):
    # BASELINE.json config 1: K=1024, T=100, dt=0.01, no smoothing
    run_pair(oracle, abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective(), 1024, 1.0, np.zeros(4), 6, 0.05, smoothing=None)


def test_toy_keep_best_and_smoothing(oracle):
    run_pair(oracle, abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective(), 300, 0.5, np.array([0.1, -0.2, 0.3, 0.0]), 8, 0.05, keep=20)


def test_toy_odd_cadence(oracle):
    # sub-step residue accumulates until a whole step is reached (SURVEY A-2)
    run_pair(oracle, abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, abi.default_toy_objective(), 64, 0.3, np.zeros(4), 10, 0.013, keep=20)


@pytest.mark.parametrize("mode", [abi.DYNAMICS_FUSED, abi.DYNAMICS_FAITHFUL])
def test_franka_trackpoint_fp64(oracle, mode):
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 512, 0.64, abi.huddled_state(), 3, 0.05, mode=mode)


def test_franka_trackpoint_config2_full_size(oracle):
    # BASELINE.json config 2: K=4096 x T=64, FP64, identical injected noise (one warm update + one shifted)
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 4096, 0.64, abi.huddled_state(), 2, 0.05, keep=20)


def test_franka_trackpoint_all_terms(oracle):
    tp = abi.default_track_point()
    tp.enable_self_collision_avoidance, tp.enable_reach_limits, tp.link_position_mode = 1, 1, abi.LINKS_BODY_COM
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, tp, 128, 0.3, abi.huddled_state(), 3, 0.05, keep=20)


@pytest.mark.parametrize("energy,links", [(True, abi.LINKS_BODY_COM), (False, abi.LINKS_ZERO)])
@pytest.mark.parametrize("mode", [abi.DYNAMICS_FUSED, abi.DYNAMICS_FAITHFUL])
def test_franka_assisted_fp64(oracle, energy, links, mode):
    x0 = abi.huddled_state(10.0)
    x0[12:24] = 0.02
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(energy, links), 256, 0.32, x0, 3, 0.05,
             keep=20, wrench=cases.constant_wrench(32), mode=mode)


def test_franka_assisted_no_forecast(oracle):
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), 128, 0.3,
             abi.huddled_state(10.0), 2, 0.05, wrench=None)


def test_franka_assisted_fp32_fast_mode(oracle):
    # BASELINE.json config 3 tolerance: updated control sequence within 1e-4 relative of the FP64 oracle
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, cases.assisted_params(True, abi.LINKS_BODY_COM), 1024, 0.64,
             abi.huddled_state(10.0), 3, 0.05, wrench=cases.constant_wrench(64), precision=abi.FP32, cost_rtol=5e-4, u_rtol=U_RTOL_FP32)


def test_franka_trackpoint_fp32_fast_mode(oracle):
    run_pair(oracle, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, abi.default_track_point(), 1024, 0.64, abi.huddled_state(), 3, 0.05,
             precision=abi.FP32, cost_rtol=5e-4, u_rtol=U_RTOL_FP32)


def test_nan_rollouts_get_zero_weight(oracle):
    # a NaN in one rollout's noise makes its cost NaN -> weight 0 (mppi.cpp:331-334,385-388)
    import engine_lib as el
    K, T = 62, 20
    holder = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.2, smoothing=None)
    o, e = ol.Oracle(oracle, holder, abi.default_toy_objective()), el.Engine(holder, abi.default_toy_objective())
    eps = np.random.default_rng(3).standard_normal((K + 2, T, 2))
    eps[5, 3, 0] = np.nan
    eps[40, 0, 1] = np.nan
    assert o.update(np.zeros(4), 0.0, None, eps) == 0 and e.update(np.zeros(4), 0.0, None, eps) == 0
    co, ce = o.read(abi.READ_COSTS, K + 2), e.read(abi.READ_COSTS, K + 2)
    assert np.isnan(ce[5]) and np.isnan(ce[40]) and np.array_equal(np.isnan(co), np.isnan(ce))
    we = e.read(abi.READ_WEIGHTS, K + 2)
    assert we[5] == 0.0 and we[40] == 0.0 and abs(we.sum() - 1.0) < 1e-12
    # 0 * NaN = NaN in the weighted sum (mppi.cpp:415-418): the poisoned entries are NaN on both sides
    Uo, Ue = o.read(abi.READ_OPTIMAL, 2 * T), e.read(abi.READ_OPTIMAL, 2 * T)
    assert np.array_equal(np.isnan(Uo), np.isnan(Ue)) and np.isnan(Uo).sum() == 2
    assert np.allclose(Ue[~np.isnan(Uo)], Uo[~np.isnan(Uo)], rtol=1e-9, atol=1e-12)
    # all NaN -> error, nothing published (mppi.cpp:368-370)
    before = e.read(abi.READ_OPTIMAL, 2 * T)
    assert e.update(np.array([np.nan, 0, 0, 0]), 0.05, None, eps) == abi.ERR_ALL_NAN
    assert "all nan rollouts" in e.error()
    assert np.array_equal(e.read(abi.READ_OPTIMAL, 2 * T), before, equal_nan=True)
    o.close()
    e.close()


def test_equal_costs_early_return(oracle):
    # max - min < 1e-6 -> weights, gradient and U untouched, no smoothing, no clamp (mppi.cpp:373-375)
    import engine_lib as el
    K, T = 30, 10
    holder = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, K, 0.1)
    p = abi.default_toy_objective()
    p.position_cost = p.velocity_cost = p.control_cost = 0.0
    o, e = ol.Oracle(oracle, holder, p), el.Engine(holder, p)
    eps = np.random.default_rng(4).standard_normal((K + 2, T, 2))
    assert o.update(np.zeros(4), 0.0, None, eps) == 0 and e.update(np.zeros(4), 0.0, None, eps) == 0
    assert np.array_equal(e.read(abi.READ_OPTIMAL, 2 * T), o.read(abi.READ_OPTIMAL, 2 * T))
    assert np.all(e.read(abi.READ_WEIGHTS, K + 2) == 0.0)
    o.close()
    e.close()


def test_create_error_conventions():
    # mppi.cpp:18-69: nullptr + reason
    import ctypes as C
    import engine_lib as el
    p = abi.default_toy_objective()

    def rc(**kw):
        system = kw.pop("system", abi.SYSTEM_TOY)
        objective = kw.pop("objective", abi.OBJECTIVE_TOY)
        h = abi.make_config(system, objective, kw.pop("K", 8), 1.0, **kw)
        out = C.c_void_p()
        code = el.lib().mppi_b200_create(C.byref(h.cfg), C.cast(C.byref(p), C.c_void_p), C.sizeof(p), C.byref(out))
        if code == 0:
            el.lib().mppi_b200_destroy(out)
        return code, el.lib().mppi_b200_last_error(None).decode()
    assert rc()[0] == 0
    assert rc(K=0) == (abi.ERR_INVALID, "trajectory rollouts must be greater than zero")
    assert rc(keep_best=-1) == (abi.ERR_INVALID, "trajectory cached rollouts cannot be less than zero")
    assert rc(threads=0) == (abi.ERR_INVALID, "trajectory threads must be positive nonzero")
    assert rc(control_min=np.zeros(3))[1] == "controller maximum and minimum must have length 2"
    assert rc(covariance=np.eye(3))[1].startswith("controller sample variance dof 3")
    assert rc(covariance=np.ones((2, 3)))[1] == "controller covariance matrix not square"
    assert rc(system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_TOY)[0] == abi.ERR_UNSUPPORTED


def test_time_must_be_monotonic_for_smoothing():
    import engine_lib as el
    holder = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 16, 0.2)
    e = el.Engine(holder, abi.default_toy_objective())
    assert e.update(np.zeros(4), 0.0) == 0
    assert e.update(np.zeros(4), 0.1) == 0
    assert e.update(np.zeros(4), 0.05) == abi.ERR_TIME  # filter.cpp:37-44
    assert "Resetting the window back in the past" in e.error()
    e.close()
