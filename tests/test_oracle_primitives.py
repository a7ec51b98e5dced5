"""Known answers and self-consistency for the oracle's leaves: cost functors (reference
src/controller/cost.hpp), energy tank (energy.hpp), and the rigid-body arithmetic that stands in
for pinocchio 2.7.1 (PARITY UNPINNED there: pinocchio is absent, so these are derived checks —
SURVEY §8c(5) — not reference vectors)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from assistedmanipulation_b200 import abi

GOLD = json.load(open(os.path.join(ol.ROOT, "tests", "golden", "robot_model.json")))


def test_cost_functor_point_values(oracle):
    L, R, Q = oracle.oracle_left_barrier, oracle.oracle_right_barrier, oracle.oracle_quadratic
    assert L(0.0, 1.0, 1e10, -0.85) == 1e10 + 0.85 ** 2           # cost.hpp:90-91
    assert L(0.0, 1.0, 1e10, 0.5) == 2.0                           # scale/(value-bound)
    assert L(0.0, 1.0, 1e10, 1e-12) == 1e10                        # clamped to maximum_cost
    assert L(-2.0, 0.0, 1e10, 0.3) == 0.0                          # zero scale inside
    assert L(-2.0, 0.0, 1e10, -2.0) == 1e10                        # zero scale at the bound still costs
    assert R(1.0, 1.0, 1e10, 0.5) == 2.0
    assert R(1.0, 10.0, 1e10, 1.5) == 1e10 + 10.0 * 0.25
    assert Q(100.0, 0.0, 500.0, 0.1) == 100.0 + 0.0 * 0.1 + 500.0 * 0.1 * 0.1
    assert Q(1.0, 2.0, 3.0, -2.0) == 1.0 + 4.0 + 12.0
    assert oracle.oracle_upper_log_barrier(1.0, 2.0, 0.0, 1e10, 1.0) == 1e10      # cost.hpp:126-127
    assert oracle.oracle_upper_log_barrier(1.0, 2.0, 0.0, 1e10, 0.9) == min(2.0 * -np.log10(0.1) * 1.0, 0.0)
    assert oracle.oracle_lower_log_barrier(0.0, 1.0, -1.0, 1e10, 10.0) == min(1.0 * (-1.0 - 1.0), 0.0)


def test_energy_tank_probe(oracle):
    # 10 J, -1 W for 0.5 s -> 9.5 J, state sqrt(19)
    out = np.zeros(2)
    oracle.oracle_tank(10.0, -1.0, 0.5, ol.ptr(out))
    assert out[0] == 9.5 and out[1] == np.sqrt(19.0)
    oracle.oracle_tank(0.1, -1.0, 0.5, ol.ptr(out))
    assert out[0] == 0.0 and out[1] == 0.0  # clamped at zero (energy.hpp:20)


def fk(oracle, q):
    ee, mt, lc = np.zeros(3), np.zeros(3), np.zeros(39)
    oracle.oracle_robot_fk(ol.ptr(np.ascontiguousarray(q, dtype=float)), ol.ptr(ee), ol.ptr(mt), ol.ptr(lc))
    return ee, mt, lc.reshape(13, 3)


@pytest.mark.parametrize("preset", sorted(GOLD["fk_kat"]))
def test_fk_known_answers(oracle, preset):
    k = GOLD["fk_kat"][preset]
    ee, mt, lc = fk(oracle, k["q"])
    assert np.allclose(ee, k["frames"]["panda_grasp_joint"], rtol=0, atol=1e-14)
    assert np.allclose(mt, k["frames"]["arm_mount_joint"], rtol=0, atol=1e-14)
    for i, name in enumerate(["pivot"] + ["panda_link%d" % j for j in range(1, 8)]):
        assert np.allclose(lc[3 + i], k["link_com"][name], rtol=0, atol=1e-14)


def test_fk_survey_appendix_b(oracle):
    ee, mt, _ = fk(oracle, GOLD["fk_kat"]["HUDDLED"]["q"])
    assert np.allclose(ee, [0.870297769478, 0.877368837289, 0.890776004667], atol=1e-11)
    assert np.allclose(mt, [0.405060966544, 0.412132034356, 0.725], atol=1e-11)


def test_aba_inverts_crba(oracle):
    rng = np.random.default_rng(0)
    for _ in range(10):
        q, v, tau = rng.uniform(-1.5, 1.5, 12), rng.uniform(-1, 1, 12), rng.uniform(-10, 10, 12)
        nle, a, M = np.zeros(12), np.zeros(12), np.zeros(144)
        oracle.oracle_robot_nle(ol.ptr(q), ol.ptr(v), ol.ptr(nle))
        oracle.oracle_robot_aba(ol.ptr(q), ol.ptr(v), ol.ptr(tau), ol.ptr(a))
        oracle.oracle_robot_crba(ol.ptr(q), ol.ptr(M))
        M = M.reshape(12, 12)
        assert np.abs(M - M.T).max() == 0.0 and np.linalg.eigvalsh(M).min() > 0
        assert np.allclose(M @ a, tau - nle, rtol=1e-10, atol=1e-10)


def test_gravity_torque_is_potential_gradient(oracle):
    rng = np.random.default_rng(1)
    mass = [j["mass"] for j in GOLD["joints"]]

    def U(q):
        _, _, lc = fk(oracle, q)
        return sum(9.81 * mass[j] * lc[j + 1, 2] for j in range(12))
    q = rng.uniform(-1, 1, 12)
    g = np.zeros(12)
    oracle.oracle_robot_nle(ol.ptr(q), ol.ptr(np.zeros(12)), ol.ptr(g))
    num = np.array([(U(q + 1e-6 * e) - U(q - 1e-6 * e)) / 2e-6 for e in np.eye(12)])
    assert np.allclose(g, num, atol=1e-6)


def kin(oracle, q, v):
    pos, lin, ang, J = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(72)
    oracle.oracle_robot_kinematics(ol.ptr(np.ascontiguousarray(q)), ol.ptr(np.ascontiguousarray(v)), ol.ptr(pos), ol.ptr(lin), ol.ptr(ang), ol.ptr(J))
    return pos, lin, ang, J.reshape(6, 12)


def test_world_jacobian_against_finite_differences(oracle):
    # WORLD-frame spatial jacobian: point velocity = J_lin qd + (J_ang qd) x p  (SURVEY A-3)
    rng = np.random.default_rng(2)
    q, v = rng.uniform(-1, 1, 12), rng.uniform(-1, 1, 12)
    pos, lin, ang, J = kin(oracle, q, v)
    J = J.copy()
    # undo the reference's overwrite of the top-left 3x3 (pinocchio_dynamics.cpp:196-200) for the check
    yaw = q[2]
    assert np.allclose(J[:3, :3], [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    eps = 1e-7
    pd = (fk(oracle, q + eps * v)[0] - fk(oracle, q - eps * v)[0]) / (2 * eps)
    # spatial velocity at the world origin reported by the reference: (lin, ang); point velocity follows
    assert np.allclose(lin + np.cross(ang, pos), pd, atol=1e-6)
    # columns 3..9 (the ones manipulability reads) reproduce the spatial velocity of the arm joints
    vv = np.zeros(12)
    vv[3:10] = v[3:10]
    _, lin2, ang2, _ = kin(oracle, q, vv)
    assert np.allclose(J[:3, 3:10] @ v[3:10], lin2, atol=1e-12)
    assert np.allclose(J[3:, 3:10] @ v[3:10], ang2, atol=1e-12)
    assert np.all(J[:, 10:] == 0.0)  # fingers do not support the grasp frame


def test_energy_is_conserved_without_control(oracle):
    # u = 0 -> tau = nle exactly cancels: qdd = 0 and the arm coasts (M^-1 * 0); kinetic energy via CRBA
    x = abi.huddled_state()
    x[12 + 3:12 + 10] = 0.1
    n = 50
    out = np.zeros((n, 31))
    oracle.oracle_robot_rollout(ol.ptr(x), ol.ptr(np.zeros((n, 12))), n, 0.001, ol.ptr(out))
    assert np.allclose(out[:, 12 + 3:12 + 10], 0.1, atol=1e-9)  # zero net generalized force: velocities unchanged
    assert np.allclose(out[-1, 3:10], x[3:10] + 0.1 * 0.001 * n, atol=1e-9)


def test_op_count_per_rollout_step(oracle):
    tp = abi.default_track_point()
    cnt = (C.c_uint64 * 6)()
    n = oracle.oracle_count_step_flops(abi.OBJECTIVE_TRACK_POINT, C.cast(C.byref(tp), C.c_void_p), cnt)
    assert 15000 < n < 30000 and sum(cnt[:5]) == n


# ---- cross-check against a derivation that shares no formulation with the oracle (tests/independent_dynamics.py) ----

def test_mass_matrix_against_textbook_jacobian_form(oracle):
    import independent_dynamics as ind
    rng = np.random.default_rng(5)
    for _ in range(6):
        q = rng.uniform(-1.5, 1.5, 12)
        M = np.zeros(144)
        oracle.oracle_robot_crba(ol.ptr(q), ol.ptr(M))
        Mi, _ = ind.mass_matrix_and_potential(q)
        assert np.abs(M.reshape(12, 12) - Mi).max() <= 1e-11 * np.abs(Mi).max()
        assert np.allclose(fk(oracle, q)[0], ind.end_effector_position(q), atol=1e-13)


def test_nonlinear_effects_against_the_lagrangian(oracle):
    """RNEA's Coriolis/centrifugal + gravity vector equals Mdot v - 1/2 d(v^T M v)/dq + dU/dq of the textbook M, U."""
    import independent_dynamics as ind
    rng = np.random.default_rng(6)
    for _ in range(4):
        q, v = rng.uniform(-1.5, 1.5, 12), rng.uniform(-2, 2, 12)
        nle = np.zeros(12)
        oracle.oracle_robot_nle(ol.ptr(q), ol.ptr(v), ol.ptr(nle))
        want = ind.nonlinear_effects(q, v)
        assert np.abs(nle - want).max() <= 2e-7 * max(1.0, np.abs(want).max()), (nle, want)


def test_forward_dynamics_against_the_lagrangian(oracle):
    """ABA(q, v, tau) = M^-1 (tau - C v - g) with M, C, g from the independent derivation."""
    import independent_dynamics as ind
    rng = np.random.default_rng(7)
    for _ in range(3):
        q, v, tau = rng.uniform(-1.5, 1.5, 12), rng.uniform(-1, 1, 12), rng.uniform(-10, 10, 12)
        a = np.zeros(12)
        oracle.oracle_robot_aba(ol.ptr(q), ol.ptr(v), ol.ptr(tau), ol.ptr(a))
        M, _ = ind.mass_matrix_and_potential(q)
        want = np.linalg.solve(M, tau - ind.nonlinear_effects(q, v))
        assert np.abs(a - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), (a, want)
