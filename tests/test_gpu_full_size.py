"""Parity at BASELINE.json's full sizes. Where the oracle finishes in seconds it is compared directly
(configs 2 and 3, and the K = 65 536 subset of config 4); at K = 1 048 576 the checks are the
size-independent properties of the path: reproducibility, argmin = lowest index of the minimum cost,
weights that sum to one, the weighted sum recomputed from the read-back noise, bounded controls."""
import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_config3_full_size_fp32_fast_mode(oracle, seed):
    """BASELINE.json config 3 at full size — K = 16384 x T = 128, AssistedManipulation + energy tank + body-COM links +
    forecast table — FP32 fast mode against the FP64 oracle on the noise the kernel consumed. The north-star gate: the
    updated control sequence within 1e-4 relative. The objective has 1e10 steps (cost.hpp:59-61,90-92); the fast mode
    carries the whole state path in FP64 (rollout_core.cuh MIXED_SOLVER) so that a rollout takes them where its FP64 twin
    does: measured 0 / 1 / 3 rollouts of 16386 on the other side for these seeds (single-precision link distances),
    control sequence 5e-6 / 2e-5 / 5e-5 (round 1, all single precision: 9..15 rollouts, 1e-4..2e-4)."""
    import engine_lib as el
    K, T, nu = 16384, 128, 12
    params, W, x0 = cases.assisted_params(True, abi.LINKS_BODY_COM), cases.constant_wrench(T), abi.huddled_state(10.0)
    e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 1.28, precision=abi.FP32, dynamics_mode=abi.DYNAMICS_FUSED,
                                  smoothing=None, control_bound=False), params)
    o = ol.Oracle(oracle, abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 1.28, threads=16, smoothing=None, control_bound=False), params)
    assert e.update(x0, 0.0, W, seed=seed) == 0, e.error()
    noise = e.read(abi.READ_NOISE, (K + 2) * T * nu)          # FP32 values widened: exactly what the kernel consumed
    assert o.update(x0, 0.0, W, noise) == 0
    Uo, Ue = o.read(abi.READ_OPTIMAL, nu * T), e.read(abi.READ_OPTIMAL, nu * T)
    co, ce = o.read(abi.READ_COSTS, K + 2), e.read(abi.READ_COSTS, K + 2)
    flipped = np.abs(ce - co) > 5e9
    assert flipped.sum() <= 6, flipped.sum()
    assert np.median(np.abs(ce - co) / np.abs(co)) < 1e-6
    assert (np.abs(ce - co) / np.abs(co))[~flipped].max() < 1e-3
    # THE GATE (BASELINE.json north_star): updated control sequence within 1e-4 relative in the FP32 fast mode
    assert np.abs(Ue - Uo).max() <= 1e-4 * np.abs(Uo).max(), np.abs(Ue - Uo).max() / np.abs(Uo).max()
    e.close()
    o.close()


def test_config4_subset_65536_fp64(oracle):
    # BASELINE.json config 4 parity subset: K = 65 534 (+2), T = 64, FP64, Philox noise read back and injected into the oracle
    import engine_lib as el
    K, T, nu = 65534, 64, 12
    tp, x0 = abi.default_track_point(), abi.huddled_state()
    e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.64, dynamics_mode=abi.DYNAMICS_FUSED), tp)
    o = ol.Oracle(oracle, abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.64, threads=16), tp)
    for u in range(2):
        assert e.update(x0, 0.05 * u, seed=9) == 0, e.error()
        noise = e.read(abi.READ_NOISE, (K + 2) * T * nu)
        assert o.update(x0, 0.05 * u, None, noise) == 0
        co, ce = o.read(abi.READ_COSTS, K + 2), e.read(abi.READ_COSTS, K + 2)
        assert (np.abs(ce - co) / np.abs(co)).max() <= 1e-9
        assert e.query(abi.QUERY_ARGMIN) == o.query(abi.QUERY_ARGMIN)
        Uo, Ue = o.read(abi.READ_OPTIMAL, nu * T), e.read(abi.READ_OPTIMAL, nu * T)
        assert np.abs(Ue - Uo).max() <= 1e-9 * np.abs(Uo).max()
    e.close()
    o.close()


@pytest.mark.parametrize("precision", [abi.FP32, abi.FP64])
def test_config4_full_size_properties(precision):
    import engine_lib as el
    K, T, nu = 1048576, 64, 12
    tp, x0 = abi.default_track_point(), abi.huddled_state()
    mk = lambda: abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.64, precision=precision, dynamics_mode=abi.DYNAMICS_FUSED)
    e = el.Engine(mk(), tp)
    Us = []
    for u in range(2):
        assert e.update(x0, 0.05 * u, seed=21) == 0, e.error()
        Us.append(e.read(abi.READ_OPTIMAL, nu * T))
    costs, w = e.read(abi.READ_COSTS, K + 2), e.read(abi.READ_WEIGHTS, K + 2)
    assert np.all(np.isfinite(costs))
    assert e.query(abi.QUERY_ARGMIN) == int(np.argmin(costs))                      # np.argmin = first index of the minimum
    mm = e.read(abi.READ_MINMAX, 2)
    assert mm[0] == costs.min() and mm[1] == costs.max()
    assert abs(w.sum() - 1.0) < 1e-9 and w.min() >= 0.0
    # the weights are the reference's formula evaluated on the read-back costs (mppi.cpp:391-393)
    ref_w = np.exp(-10.0 * (costs - mm[0]) / (mm[1] - mm[0]))
    assert np.allclose(w, ref_w / ref_w.sum(), rtol=1e-9, atol=1e-18)
    # bounded controls (mppi.cpp:443-447)
    U = Us[-1].reshape(T, nu)
    assert np.all(U <= abi.FRANKA_CONTROL_MAX + 1e-15) and np.all(U >= abi.FRANKA_CONTROL_MIN - 1e-15)
    # reproducible: a second engine with the same seed publishes the same bits
    e2 = el.Engine(mk(), tp)
    for u in range(2):
        assert e2.update(x0, 0.05 * u, seed=21) == 0
        assert np.array_equal(e2.read(abi.READ_OPTIMAL, nu * T), Us[u])
    e.close()
    e2.close()


def test_weighted_sum_recomputed_from_read_back_noise():
    # linearity check of K4 at K = 131 070: gradient == (w^T eps) / sum(w) recomputed in numpy from what the device holds
    import engine_lib as el
    K, T, nu = 131070, 64, 12
    e = el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.64, dynamics_mode=abi.DYNAMICS_FUSED, smoothing=None, control_bound=False), abi.default_track_point())
    x0 = abi.huddled_state()
    assert e.update(x0, 0.0, seed=5) == 0
    U0 = e.read(abi.READ_OPTIMAL, nu * T)
    assert e.update(x0, 0.05, seed=5) == 0
    w = e.read(abi.READ_WEIGHTS, K + 2)
    noise = e.read(abi.READ_NOISE, (K + 2) * T * nu).reshape(K + 2, T * nu)
    g = e.read(abi.READ_GRADIENT, nu * T)
    assert np.allclose(g, w @ noise, rtol=1e-10, atol=1e-13)
    # and the update is the shifted previous optimum plus gradient_step * gradient (mppi.cpp:204-206,421)
    U0 = U0.reshape(T, nu)
    shifted = np.vstack([U0[5:], np.repeat(U0[-1:], 5, axis=0)]).reshape(-1)
    assert np.allclose(e.read(abi.READ_OPTIMAL, nu * T), shifted + 2.0 * g, rtol=1e-12, atol=1e-14)
    e.close()
