"""The engine's host/device templates (csrc/robot.cuh, rollout_core.cuh) compiled for the CPU and
checked against the oracle — an independent implementation (the device code is specialised by joint
type, carries the articulated inertia in symmetric blocks and has a FUSED evaluation mode; the
oracle follows pinocchio's generic recursion). This is test infrastructure: the product library has
no CPU path (tests/host_check builds its own shared object)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi

HC_DIR = os.path.join(ol.ROOT, "tests", "host_check")
_dp = ol._dp


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC_DIR, "libhost_check.so")
    src = os.path.join(HC_DIR, "host_check.cpp")
    csrc = os.path.join(ol.ROOT, "assistedmanipulation_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("robot.cuh", "robot_fast.cuh", "rollout_core.cuh", "spatial.cuh", "model_init.h", "params_convert.h", "robot_model.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-x", "c++", "-fPIC", "-shared", "-I" + csrc, "-o", so, src])
    lib = C.CDLL(so)
    lib.host_robot_calculate.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.host_rollouts.argtypes = [C.c_int, C.c_int, C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp]
    return lib


def test_device_sincos_matches_libm(hc):
    """robot_fast.cuh sincos_poly (magic-constant reduction, sign bits flipped on the integer side) against libm: every
    quadrant, both signs, exact zero, the neighbourhood of multiples of pi/2 and arguments up to the 2^31 range limit."""
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e4, 1e4, 50000), rng.uniform(-2.0e9, 2.0e9, 50000),
                        np.arange(-64, 65) * (np.pi / 2), np.arange(-64, 65) * (np.pi / 2) + 1e-9, np.arange(-64, 65) * (np.pi / 4),
                        [0.0, -0.0, 1e-300, -1e-300, 5e-324, 0.7853981633974483, -0.7853981633974483]])
    s, c = np.zeros_like(a), np.zeros_like(a)
    hc.host_sincos_poly.argtypes = [C.c_int, _dp, _dp, _dp]
    hc.host_sincos_poly(a.size, ol.ptr(a), ol.ptr(s), ol.ptr(c))
    assert np.abs(s - np.sin(a)).max() <= 2.3e-16 and np.abs(c - np.cos(a)).max() <= 2.3e-16     # one ulp of a value near 1
    small = np.abs(np.sin(a)) < 0.5
    assert (np.abs(s - np.sin(a))[small] <= 2.0 * np.spacing(np.abs(np.sin(a))[small]) + 1e-25).all()   # relative accuracy near the zeros
    assert (s[a == 0.0] == 0.0).all() and (c[a == 0.0] == 1.0).all()
    assert np.abs(s * s + c * c - 1.0).max() <= 5e-16


def test_device_sincos_f32_matches_libm(hc):
    rng = np.random.default_rng(4)
    a = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-3000, 3000, 50000), rng.uniform(-1e5, 1e5, 50000), np.arange(-64, 65) * (np.pi / 4), [0.0, -0.0]]).astype(np.float32)
    s, c = np.zeros_like(a), np.zeros_like(a)
    fp = C.POINTER(C.c_float)
    hc.host_sincos_poly_f32.argtypes = [C.c_int, fp, fp, fp]
    hc.host_sincos_poly_f32(a.size, a.ctypes.data_as(fp), s.ctypes.data_as(fp), c.ctypes.data_as(fp))
    assert np.abs(s - np.sin(a.astype(np.float64))).max() <= 1.2e-7 and np.abs(c - np.cos(a.astype(np.float64))).max() <= 1.2e-7
    assert (s[a == 0.0] == 0.0).all() and (c[a == 0.0] == 1.0).all()
    bad = np.array([np.inf, -np.inf, np.nan], np.float32)
    s3, c3 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    hc.host_sincos_poly_f32(3, bad.ctypes.data_as(fp), s3.ctypes.data_as(fp), c3.ctypes.data_as(fp))
    assert np.isnan(s3).all() and np.isnan(c3).all()


def test_integer_sign_test_matches_the_comparison(hc):
    """is_negative(x) (sign bit set and not a NaN, decided on the integer pipe) has the truth value of the reference's
    `x < 0` for every non-zero x, NaN of either sign included (false); -0, which a difference of finite numbers never
    produces, counts as negative."""
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.standard_normal(10000) * 10.0 ** rng.integers(-300, 300, 10000), [np.inf, -np.inf, 5e-324, -5e-324, 1e-40, -1e-40, 0.0,
                        np.nan, -np.nan, np.float64(np.frombuffer(np.uint64(0xfff8000000000001).tobytes(), np.float64)[0]), np.float64(np.frombuffer(np.uint64(0x7ff8000000000001).tobytes(), np.float64)[0])]])
    o64, o32 = np.zeros(x.size, np.int32), np.zeros(x.size, np.int32)
    ip = C.POINTER(C.c_int)
    hc.host_is_negative.argtypes = [C.c_int, _dp, ip, ip]
    hc.host_is_negative(x.size, ol.ptr(x), o64.ctypes.data_as(ip), o32.ctypes.data_as(ip))
    assert np.array_equal(o64.astype(bool), x < 0)
    with np.errstate(over="ignore", under="ignore"):
        xf = x.astype(np.float32)
    nz = xf != 0                                 # values that underflow to -0 in single precision count as negative there
    assert np.array_equal(o32.astype(bool)[nz], (xf < 0)[nz])
    m0 = np.array([-0.0])
    z64, z32 = np.zeros(1, np.int32), np.zeros(1, np.int32)
    hc.host_is_negative(1, ol.ptr(m0), z64.ctypes.data_as(ip), z32.ctypes.data_as(ip))
    assert z64[0] == 1 and z32[0] == 1


def test_topology_matches_generated_model(hc):
    assert hc.host_topology_matches() == 1


@pytest.mark.parametrize("mode,tol", [(0, 1e-12), (1, 1e-12), (2, 1e-12), (4, 5e-5), (5, 5e-5)])
def test_calculate_matches_oracle(oracle, hc, mode, tol):
    rng = np.random.default_rng(0)
    for _ in range(5):
        q, v, u = rng.uniform(-1.5, 1.5, 12), rng.uniform(-1, 1, 12), np.zeros(12)
        u[3:10] = rng.uniform(-10, 10, 7)
        nle, a = np.zeros(12), np.zeros(12)
        oracle.oracle_robot_nle(ol.ptr(q), ol.ptr(v), ol.ptr(nle))
        tau = u + nle
        oracle.oracle_robot_aba(ol.ptr(q), ol.ptr(v), ol.ptr(tau), ol.ptr(a))
        pos, lin, ang, J = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(72)
        oracle.oracle_robot_kinematics(ol.ptr(q), ol.ptr(v), ol.ptr(pos), ol.ptr(lin), ol.ptr(ang), ol.ptr(J))
        Ja = J.reshape(6, 12)[:3, 3:10]
        ee, mt, lc = np.zeros(3), np.zeros(3), np.zeros(39)
        oracle.oracle_robot_fk(ol.ptr(q), ol.ptr(ee), ol.ptr(mt), ol.ptr(lc))
        qdd, n2, kin = np.zeros(12), np.zeros(12), np.zeros(34)
        hc.host_robot_calculate(mode, ol.ptr(q), ol.ptr(v), ol.ptr(u), ol.ptr(qdd), ol.ptr(n2), ol.ptr(kin))
        assert np.abs(qdd - a).max() <= tol * np.abs(a).max()
        if mode & 3:
            assert np.abs(n2 - nle).max() <= tol * max(np.abs(nle).max(), 1.0)
        assert np.abs(kin[:3] - pos).max() <= tol and np.abs(kin[3:6] - mt).max() <= tol
        assert np.abs(kin[6:9] - lin).max() <= tol * 10
        assert abs(kin[9] - np.linalg.det(Ja @ Ja.T)) <= max(tol, 1e-10) * 10 * abs(np.linalg.det(Ja @ Ja.T)) + tol * 1e-3
        assert np.abs(kin[10:].reshape(8, 3) - lc.reshape(13, 3)[3:11]).max() <= tol


def _oracle_costs(oracle, objective, params, K, T, x0, wrench, seed):
    holder = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, objective, K, T * 0.01, smoothing=None, threads=4)
    o = ol.Oracle(oracle, holder, params)
    eps = np.random.default_rng(seed).standard_normal((K + 2, T, 12)) * np.sqrt(abi.FRANKA_COVARIANCE_DIAG)
    assert o.update(x0, 0.0, wrench, eps) == 0
    costs, noise = o.read(abi.READ_COSTS, K + 2), o.read(abi.READ_NOISE, (K + 2) * T * 12)
    bd = o.read(abi.READ_BREAKDOWN, 8)
    U = o.read(abi.READ_OPTIMAL, 12 * T)
    o.close()
    return costs, noise, bd, U


CASES = {
    "track_point": (abi.OBJECTIVE_TRACK_POINT, abi.default_track_point, None),
    "assisted_energy_com": (abi.OBJECTIVE_ASSISTED_MANIPULATION, lambda: cases.assisted_params(True, abi.LINKS_BODY_COM), cases.constant_wrench(32)),
    "assisted_zero_links": (abi.OBJECTIVE_ASSISTED_MANIPULATION, lambda: cases.assisted_params(False, abi.LINKS_ZERO), cases.constant_wrench(32)),
    "assisted_no_forecast": (abi.OBJECTIVE_ASSISTED_MANIPULATION, lambda: cases.assisted_params(True, abi.LINKS_BODY_COM), None),
}


@pytest.mark.parametrize("f32,tol", [(0, 1e-12), (1, 5e-5)])
def test_fused_step_by_joint_structure_matches_oracle(oracle, hc, f32, tol):
    """What the FUSED rollout step executes — kinematics / RNEA crossing joints 0..9 by their structure (PLANE, robot.cuh)
    on the shared polynomial sines / cosines, and the solver that also carries the end effector point — against the
    oracle's generic recursion: accelerations, nonlinear effects, every kinematic quantity the objectives read."""
    rng = np.random.default_rng(7)
    hc.host_robot_calculate_plane.argtypes = [C.c_int] + [_dp] * 7
    for _ in range(8):
        q, v, u = rng.uniform(-1.5, 1.5, 12), rng.uniform(-1, 1, 12), np.zeros(12)
        q[10:] = rng.uniform(0.0, 0.04, 2)
        u[3:10] = rng.uniform(-10, 10, 7)
        nle, a = np.zeros(12), np.zeros(12)
        oracle.oracle_robot_nle(ol.ptr(q), ol.ptr(v), ol.ptr(nle))
        tau = u + nle
        oracle.oracle_robot_aba(ol.ptr(q), ol.ptr(v), ol.ptr(tau), ol.ptr(a))
        pos, lin, ang, J = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(72)
        oracle.oracle_robot_kinematics(ol.ptr(q), ol.ptr(v), ol.ptr(pos), ol.ptr(lin), ol.ptr(ang), ol.ptr(J))
        Ja = J.reshape(6, 12)[:3, 3:10]
        ee, mt, lc = np.zeros(3), np.zeros(3), np.zeros(39)
        oracle.oracle_robot_fk(ol.ptr(q), ol.ptr(ee), ol.ptr(mt), ol.ptr(lc))
        qdd, n2, kin, ee3 = np.zeros(12), np.zeros(12), np.zeros(34), np.zeros(3)
        hc.host_robot_calculate_plane(f32, ol.ptr(q), ol.ptr(v), ol.ptr(u), ol.ptr(qdd), ol.ptr(n2), ol.ptr(kin), ol.ptr(ee3))
        assert np.abs(qdd - a).max() <= tol * np.abs(a).max()
        assert np.abs(n2 - nle).max() <= tol * max(np.abs(nle).max(), 1.0)
        assert np.abs(kin[:3] - pos).max() <= tol and np.abs(ee3 - pos).max() <= tol and np.abs(kin[3:6] - mt).max() <= tol
        assert np.abs(kin[6:9] - lin).max() <= tol * 10
        assert abs(kin[9] - np.linalg.det(Ja @ Ja.T)) <= max(tol, 1e-10) * 10 * abs(np.linalg.det(Ja @ Ja.T)) + tol * 1e-3
        assert np.abs(kin[10:].reshape(8, 3) - lc.reshape(13, 3)[3:11]).max() <= tol


def test_unrolled_solver_matches_loop_body_solver_and_oracle(oracle, hc):
    """aba_fused_fast with the arm joints unrolled (the build for large rollout sets) drops the products with the
    structural zeros of the joint offsets in both passes; same accelerations and end effector point as the oracle."""
    rng = np.random.default_rng(11)
    hc.host_fast_aba_unrolled.argtypes = [_dp] * 4
    hc.host_fast_aba.argtypes = [C.c_int] + [_dp] * 4
    for _ in range(8):
        q, u = rng.uniform(-1.5, 1.5, 12), np.zeros(12)
        q[10:] = rng.uniform(0.0, 0.04, 2)
        u[3:10] = rng.uniform(-10, 10, 7)
        a, a1, a7, e1, e7 = np.zeros(12), np.zeros(12), np.zeros(12), np.zeros(3), np.zeros(3)
        nle = np.zeros(12)   # the reference's tau = u + nle(q, v): the fused solver returns M^-1 u
        oracle.oracle_robot_nle(ol.ptr(q), ol.ptr(np.zeros(12)), ol.ptr(nle))
        oracle.oracle_robot_aba(ol.ptr(q), ol.ptr(np.zeros(12)), ol.ptr(u + nle), ol.ptr(a))
        hc.host_fast_aba(0, ol.ptr(q), ol.ptr(u), ol.ptr(a1), ol.ptr(e1))
        hc.host_fast_aba_unrolled(ol.ptr(q), ol.ptr(u), ol.ptr(a7), ol.ptr(e7))
        assert np.abs(a7 - a).max() <= 1e-12 * np.abs(a).max() and np.abs(a7 - a1).max() <= 1e-12 * np.abs(a).max()
        assert np.abs(e7 - e1).max() <= 1e-14
        pos, lin, ang, J = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(72)
        oracle.oracle_robot_kinematics(ol.ptr(q), ol.ptr(np.zeros(12)), ol.ptr(pos), ol.ptr(lin), ol.ptr(ang), ol.ptr(J))
        assert np.abs(e7 - pos).max() <= 1e-12


def _track_point_full():
    tp = abi.default_track_point()
    tp.enable_self_collision_avoidance, tp.enable_reach_limits, tp.link_position_mode = 1, 1, abi.LINKS_BODY_COM
    return tp


CASES["track_point_all_terms"] = (abi.OBJECTIVE_TRACK_POINT, _track_point_full, None)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("flags,tol", [(0, 1e-9), (1, 1e-9), (4, 2e-4), (8, 1e-9), (12, 2e-4)])
def test_rollout_costs_match_oracle(oracle, hc, name, flags, tol):
    objective, make, wrench = CASES[name]
    params = make()
    K, T = 48, 32
    x0 = abi.huddled_state(10.0)
    x0[12:24] = 0.05
    costs, noise, _, _ = _oracle_costs(oracle, objective, params, K, T, x0, wrench, 5)
    out = np.zeros(K + 2)
    hc.host_rollouts(objective, flags, C.cast(C.byref(params), C.c_void_p), ol.ptr(x0), ol.ptr(np.zeros(12 * T)), ol.ptr(wrench) if wrench is not None else None,
                     ol.ptr(noise), K + 2, T, 0.01, 1.0, ol.ptr(out), None)
    assert (np.abs(out - costs) / np.abs(costs)).max() <= tol
    if flags < 4:
        assert out.argmin() == costs.argmin()


def test_optimal_breakdown_matches_oracle(oracle, hc):
    # per-term totals of the optimal re-rollout (logging/assisted_manipulation.cpp:58-103)
    objective, make, wrench = CASES["assisted_energy_com"]
    params = make()
    K, T = 32, 32
    x0 = abi.huddled_state(10.0)
    _, _, bd, U = _oracle_costs(oracle, objective, params, K, T, x0, wrench, 9)
    out, got = np.zeros(1), np.zeros(7)
    hc.host_rollouts(objective, 0, C.cast(C.byref(params), C.c_void_p), ol.ptr(x0), ol.ptr(U), ol.ptr(wrench), ol.ptr(np.zeros(12 * T)), 1, T, 0.01, 1.0, ol.ptr(out), ol.ptr(got))
    assert np.allclose(got, bd[:7], rtol=1e-9, atol=1e-9)
    assert abs(out[0] - bd[7]) <= 1e-9 * abs(bd[7])


@pytest.mark.parametrize("name", ["quadratic", "left_inverse", "left_inverse_zero_scale", "right_inverse", "right_inverse_small_max", "upper_log", "lower_log"])
def test_kernel_cost_functors_match_the_reference(hc, name):
    """The cost.hpp functors as the kernels evaluate them (rollout_core.cuh: barriers as selects, the logarithmic barriers
    of cost.hpp:105-167) against values produced by the REFERENCE's own functors (tests/golden/ref_objective.npz, generated
    from cost.hpp compiled unmodified): FP64 to the last bits, FP32 to single precision away from the 1e10 steps."""
    import objective_probe as op
    golden = np.load(os.path.join(ol.ROOT, "tests", "golden", "ref_objective.npz"))
    kind, a, b, c, d = op.FUNCTORS[name]
    v = op.functor_values(kind, a)
    want = golden["functor/" + name]
    hc.host_cost_functor.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, C.c_long, _dp]
    got = np.zeros_like(v)
    hc.host_cost_functor(kind, 0, a, b, c, d, ol.ptr(v), len(v), ol.ptr(got))
    assert np.allclose(got, want, rtol=4e-16, atol=0), np.abs(got - want).max()   # selects instead of branches: the same value in every case (a division may differ in the last bit: scale / x as written)
    got32 = np.zeros_like(v)
    hc.host_cost_functor(kind, 1, a, b, c, d, ol.ptr(v), len(v), ol.ptr(got32))
    # single precision: the bound itself is rounded, so values within 1e-6 of it may sit on the other side of the step
    away = np.abs(v - a) > 1e-5 * max(abs(a), 1.0)
    assert np.allclose(got32[away], want[away], rtol=2e-5, atol=1e-6 * np.abs(want[away]).max())


@pytest.mark.parametrize("f32", [0, 1])
def test_rolled_passes_reproduce_the_unrolled_ones(hc, f32):
    """robot.cuh robot_calculate_rolled (kinematics + RNEA with the seven arm joints as ONE loop body: a third of the code,
    for the kernels bound by instruction fetch) against the unrolled structure-crossing passes: the loop multiplies through
    the structural zeros the unrolled build drops, nothing else differs — identical values."""
    rng = np.random.default_rng(23)
    hc.host_robot_calculate_rolled.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
    hc.host_robot_calculate_plane.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    for _ in range(200):
        q = abi.huddled_state()[:12] + rng.normal(0, 0.7, 12)
        qd = rng.normal(0, 1.5, 12)
        u = rng.normal(0, 3.0, 12)
        nle_r, kin_r = np.zeros(12), np.zeros(34)
        qdd, nle_u, kin_u, ee = np.zeros(12), np.zeros(12), np.zeros(34), np.zeros(3)
        hc.host_robot_calculate_rolled(f32, ol.ptr(q), ol.ptr(qd), ol.ptr(nle_r), ol.ptr(kin_r))
        hc.host_robot_calculate_plane(f32, ol.ptr(q), ol.ptr(qd), ol.ptr(u), ol.ptr(qdd), ol.ptr(nle_u), ol.ptr(kin_u), ol.ptr(ee))
        assert np.array_equal(nle_r, nle_u), np.abs(nle_r - nle_u).max()
        assert np.array_equal(kin_r, kin_u), np.abs(kin_r - kin_u).max()
