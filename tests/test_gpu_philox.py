"""In-kernel Philox noise: determinism, moments, shard invariance of the stream, and parity of a
Philox-driven update against the oracle fed with the very noise the GPU generated (read back
through MPPI_B200_READ_NOISE)."""
import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi

pytestmark = pytest.mark.gpu


def test_philox_moments_and_determinism():
    import engine_lib as el
    K, T = 4094, 64
    h = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.64)
    e1, e2 = el.Engine(h, abi.default_track_point()), el.Engine(h, abi.default_track_point())
    x0 = abi.huddled_state()
    assert e1.update(x0, 0.0, seed=1234) == 0 and e2.update(x0, 0.0, seed=1234) == 0
    n1 = e1.read(abi.READ_NOISE, (K + 2) * T * 12).reshape(K + 2, T, 12)
    n2 = e2.read(abi.READ_NOISE, (K + 2) * T * 12).reshape(K + 2, T, 12)
    assert np.array_equal(n1, n2)
    assert np.array_equal(e1.read(abi.READ_OPTIMAL, 12 * T), e2.read(abi.READ_OPTIMAL, 12 * T))  # reproducible reductions
    assert np.all(n1[0] == 0.0) and np.all(n1[1] == 0.0)  # rollout 0 = 0, rollout 1 = -U_prev = 0 at the first update
    z = n1[2:]
    std = np.sqrt(abi.FRANKA_COVARIANCE_DIAG)
    assert np.all(z[..., 10:] == 0.0)  # zero covariance on the gripper (base.hpp:82)
    for d in range(10):
        s = z[..., d] / std[d]
        assert abs(s.mean()) < 0.01 and abs(s.std() - 1.0) < 0.01
        assert abs((s ** 3).mean()) < 0.03 and abs((s ** 4).mean() - 3.0) < 0.06
    c = np.corrcoef(z[..., :10].reshape(-1, 10).T)
    assert np.abs(c - np.eye(10)).max() < 0.01
    e1.update(x0, 0.05, seed=1234)
    assert not np.array_equal(e1.read(abi.READ_NOISE, (K + 2) * T * 12).reshape(K + 2, T, 12)[2:], z)  # counter includes the update index
    e3 = el.Engine(h, abi.default_track_point())
    e3.update(x0, 0.0, seed=99)
    assert not np.array_equal(e3.read(abi.READ_NOISE, (K + 2) * T * 12).reshape(K + 2, T, 12)[2:], z)
    for e in (e1, e2, e3):
        e.close()


@pytest.mark.parametrize("precision,u_tol,c_tol", [(abi.FP64, 1e-9, 1e-9), (abi.FP32, 1e-4, 5e-4)])
def test_philox_update_matches_oracle_on_read_back_noise(oracle, precision, u_tol, c_tol):
    import engine_lib as el
    K, T = 510, 32
    ho = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, keep_best=20, threads=8)
    he = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_TRACK_POINT, K, 0.32, keep_best=20, precision=precision, dynamics_mode=abi.DYNAMICS_FUSED)
    o, e = ol.Oracle(oracle, ho, abi.default_track_point()), el.Engine(he, abi.default_track_point())
    x0 = abi.huddled_state()
    for u in range(4):
        t = 0.05 * u
        assert e.update(x0, t, seed=7) == 0
        noise = e.read(abi.READ_NOISE, (K + 2) * T * 12)
        # the oracle takes fresh columns from `injected` exactly where the engine drew fresh Philox columns
        assert o.update(x0, t, None, noise) == 0
        if precision == abi.FP64:
            no = o.read(abi.READ_NOISE, (K + 2) * T * 12).reshape(K + 2, -1)
            assert np.array_equal(no[2:], noise.reshape(K + 2, -1)[2:])
            assert e.query(abi.QUERY_ARGMIN) == o.query(abi.QUERY_ARGMIN)
        co, ce = o.read(abi.READ_COSTS, K + 2), e.read(abi.READ_COSTS, K + 2)
        if precision == abi.FP32:
            # lean reach-to-pose kernel, all single precision: its steps are the 1000-unit joint-limit penalties, worth ~1e-6
            # of the control sequence each — a handful of rollouts may take one on the other side
            cases.fp32_costs(ce, co, c_tol, max_flips=(K + 2) // 100)
        else:
            assert (np.abs(ce - co) / np.abs(co)).max() <= c_tol
        Uo, Ue = o.read(abi.READ_OPTIMAL, 12 * T), e.read(abi.READ_OPTIMAL, 12 * T)
        assert np.abs(Ue - Uo).max() <= u_tol * np.abs(Uo).max()
        if precision == abi.FP32:
            o.set_optimal(Ue)   # next update from identical inputs (see test_gpu_parity.run_pair)
    o.close()
    e.close()
