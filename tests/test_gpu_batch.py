"""Batched engine (BASELINE.json config 5): B independent controllers as ONE grid (blockIdx.y = controller)
must reproduce B separate engines bit for bit, and the oracle within the usual tolerances."""
import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu


def _inputs(B, T):
    states, wrenches = [], []
    for c in range(B):
        x0 = abi.huddled_state(10.0)
        x0[0] += 0.01 * c
        x0[2] -= 0.02 * c
        x0[12 + 4] = 0.05 * c
        states.append(x0)
        w = cases.constant_wrench(T, (10.0 - c, 2.0 * c, 0.5))
        wrenches.append(w)
    return np.stack(states), np.stack(wrenches)


@pytest.mark.parametrize("precision", [abi.FP64, abi.FP32])
def test_batch_equals_separate_engines(precision):
    import engine_lib as el
    B, K, T, nu = 3, 126, 30, 12
    params = cases.assisted_params(True, abi.LINKS_BODY_COM)
    mk = lambda batch: abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 0.3, keep_best=20, precision=precision,
                                       dynamics_mode=abi.DYNAMICS_FUSED, batch=batch)
    batched = el.Engine(mk(B), params)
    singles = [el.Engine(mk(1), params) for _ in range(B)]
    assert batched.query(abi.QUERY_BATCH) == B
    states, wrenches = _inputs(B, T)
    R = K + 2
    for u in range(4):
        t = 0.05 * u
        assert batched.update(states, t, wrenches, seed=40) == 0, batched.error()
        Ub = batched.read(abi.READ_OPTIMAL, B * nu * T).reshape(B, -1)
        cb = batched.read(abi.READ_COSTS, B * R).reshape(B, -1)
        wb = batched.read(abi.READ_WEIGHTS, B * R).reshape(B, -1)
        kb = batched.read(abi.READ_KEPT, B * 20, np.int64).reshape(B, -1)
        ob = batched.read(abi.READ_OPTIMAL_COST, B)
        gb = batched.get(t + 0.013).reshape(B, -1)
        for c, e in enumerate(singles):
            assert e.update(states[c], t, wrenches[c], seed=40 + c) == 0   # controller c draws with seed + c
            assert np.array_equal(e.read(abi.READ_OPTIMAL, nu * T), Ub[c])
            assert np.array_equal(e.read(abi.READ_COSTS, R), cb[c])
            assert np.array_equal(e.read(abi.READ_WEIGHTS, R), wb[c])
            assert np.array_equal(e.read(abi.READ_KEPT, 20, np.int64), kb[c])
            assert e.read(abi.READ_OPTIMAL_COST, 1)[0] == ob[c]
            assert np.array_equal(e.get(t + 0.013), gb[c])
    for e in singles + [batched]:
        e.close()


def test_batch_matches_oracle_on_injected_noise(oracle):
    import engine_lib as el
    B, K, T, nu = 4, 62, 20, 12
    tp = abi.default_track_point()
    batched = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.2, keep_best=10, dynamics_mode=abi.DYNAMICS_FUSED, batch=B), tp)
    oracles = [ol.Oracle(oracle, abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.2, keep_best=10, threads=2), tp) for _ in range(B)]
    states, _ = _inputs(B, T)
    rng = np.random.default_rng(8)
    for u in range(3):
        t = 0.05 * u
        eps = rng.standard_normal((B, K + 2, T, nu)) * np.sqrt(abi.FRANKA_COVARIANCE_DIAG)
        assert batched.update(states, t, None, eps) == 0, batched.error()
        Ub = batched.read(abi.READ_OPTIMAL, B * nu * T).reshape(B, -1)
        cb = batched.read(abi.READ_COSTS, B * (K + 2)).reshape(B, -1)
        for c, o in enumerate(oracles):
            assert o.update(states[c], t, None, eps[c]) == 0
            co = o.read(abi.READ_COSTS, K + 2)
            assert (np.abs(cb[c] - co) / np.abs(co)).max() <= 1e-9
            Uo = o.read(abi.READ_OPTIMAL, nu * T)
            assert np.abs(Ub[c] - Uo).max() <= 1e-9 * np.abs(Uo).max()
    batched.close()
    for o in oracles:
        o.close()


def test_batch_cannot_be_sharded():
    import ctypes as C
    import engine_lib as el
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 16, 0.2, batch=2, rank=0, world_size=2)
    p, out = abi.default_toy_objective(), C.c_void_p()
    assert el.lib().mppi_b200_create(C.byref(h.cfg), C.cast(C.byref(p), C.c_void_p), C.sizeof(p), C.byref(out)) == abi.ERR_UNSUPPORTED
