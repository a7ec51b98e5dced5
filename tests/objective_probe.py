"""Probe inputs for the objective pin: records of (state, kinematics the objective pulls out of the dynamics, tank
energy, forecast wrench) on which the REFERENCE's own objective code (objective/track_point.cpp,
objective/assisted_manipulation.cpp, cost.hpp compiled unmodified into oracle/_ref/libmppi_ref.so) and the oracle's
restatement (oracle/systems.hpp) are both evaluated. Shared by tools/gen_objective_golden.py and
tests/test_objective_reference.py. Layout of one record: see ref_objective_probe in oracle/ref_driver.cpp."""
import ctypes as C

import numpy as np

from assistedmanipulation_b200 import abi

RECORD = 159
_dp = C.POINTER(C.c_double)


def records(seed, count):
    """Random records that visit every branch of every term: joints inside / outside / exactly on their limits, links
    inside and outside the collision spheres (and coincident: the Pinocchio backend's zeros), the end effector in front
    of / behind / below the base, a zero planar offset (acos of 0/0 -> NaN, skipped), tank energy inside / below /
    above its band, forces below / above the clamp, a zero force, no forecast handle, singular and regular jacobians."""
    rng = np.random.default_rng(seed)
    r = np.zeros((count, RECORD))
    x0 = abi.huddled_state(10.0)
    for n in range(count):
        p = r[n]
        kind = n % 8
        p[:31] = x0
        p[:12] += rng.normal(0, 0.5 if kind < 6 else 3.0, 12)           # kind 6,7: well outside the joint limits
        p[12:24] = rng.normal(0, 1.0, 12)
        if kind == 1:
            p[3] = -2.8; p[9] = 2.967                                     # exactly on a lower / an upper bound
        p[31:34] = rng.normal([0.8, 0.8, 0.9], 0.6)                       # end-effector position
        p[34:37] = rng.normal(0, 0.5, 3)                                  # end-effector linear velocity
        J = rng.normal(0, 0.5, (6, 12))
        if kind == 2:
            J[:3, 3:10] = np.outer(rng.normal(size=3), rng.normal(size=7))  # rank one: det = 0 (or a tiny negative)
        if kind == 3:
            J[:3, 3:10] *= 1e4                                            # volume above the 1e5 clamp
        p[37:109] = J.reshape(-1)
        p[109:112] = rng.normal([0.4, 0.4, 0.725], 0.2)                   # ARM_MOUNT_JOINT
        links = rng.normal(0, 0.6, (13, 3))
        if kind == 4:
            links[:] = 0.0                                                # PinocchioDynamics::get_link_position
        p[112:151] = links.reshape(-1)
        p[151] = [10.0, 0.0, -1.0, 25.0, 20.0, 5.0, 1e-9, 19.999][kind]   # tank energy around [0, 20]
        p[152] = 0.0 if kind == 5 else 1.0                                # forecast handle present
        p[153:159] = rng.normal(0, 80.0, 6)
        if kind == 6:
            p[153:156] = 0.0                                              # zero force: distance 0, not > threshold
        if kind == 7:
            p[153:156] = [1e4, -1e4, 30.0]                                # clamped components
        if n % 16 == 9:
            yaw = p[2]
            p[31:33] = p[109:111] + 0.1 * np.array([np.cos(yaw), np.sin(yaw)])   # planar offset ~0 -> acos argument 0/0 or huge
    return r


def variants():
    """name -> (objective id, params)"""
    out = {}
    am = abi.default_assisted_manipulation()
    out["assisted_default"] = (abi.OBJECTIVE_ASSISTED_MANIPULATION, am)
    am = abi.default_assisted_manipulation()
    am.enable_energy_limit = 1
    am.trajectory_position_threshold = 0.3
    am.workspace_cost_yaw = abi.Quadratic(1.0, 2.0, 400.0)
    am.trajectory_velocity_cost = abi.Quadratic(3.0, 0.5, 500.0)
    out["assisted_energy_threshold"] = (abi.OBJECTIVE_ASSISTED_MANIPULATION, am)
    tp = abi.default_track_point()
    out["trackpoint_default"] = (abi.OBJECTIVE_TRACK_POINT, tp)
    tp = abi.default_track_point()
    tp.enable_self_collision_avoidance, tp.enable_reach_limits = 1, 1
    out["trackpoint_all_terms"] = (abi.OBJECTIVE_TRACK_POINT, tp)
    return out


def evaluate(fn, objective, params, recs):
    recs = np.ascontiguousarray(recs)
    out = np.zeros((len(recs), 8))
    rc = fn(objective, C.cast(C.byref(params), C.c_void_p), recs.ctypes.data_as(_dp), len(recs), out.ctypes.data_as(_dp))
    assert rc == 0
    return out


# cost.hpp functor probes: (kind, a, b, c, d)
FUNCTORS = {
    "quadratic": (0, 100.0, 3.0, 500.0, 0.0),
    "left_inverse": (1, -2.8, 10.0, 1e10, 0.0),
    "left_inverse_zero_scale": (1, -2.0, 0.0, 1e10, 0.0),
    "right_inverse": (2, 2.967, 10.0, 1e10, 0.0),
    "right_inverse_small_max": (2, 1.0, 1.0, 50.0, 0.0),
    "upper_log": (3, 1.5, 2.0, 0.25, 1e10),
    "lower_log": (4, -0.5, 3.0, -0.125, 1e10),
}


def functor_values(kind, a):
    rng = np.random.default_rng(kind + 11)
    v = np.concatenate([rng.normal(a, 2.0, 200), a + np.array([0.0, 1e-12, -1e-12, 1e-300, -1e-300, 1.0, -1.0, 1e-9, -1e-9, 1e6, -1e6])])
    return np.ascontiguousarray(v)
