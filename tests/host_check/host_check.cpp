// TEST INFRASTRUCTURE — compiles the engine's host/device templates (csrc/robot.cuh,
// objectives.cuh, rollout_core.cuh) for the CPU so their arithmetic can be checked against the
// oracle without a GPU. Built by tests/test_device_math_host.py into tests/host_check/; the
// product library (libmppi_b200.so) contains none of this and has no CPU path.
#include <cstring>
#include <vector>
#include "model_init.h"
#include "rollout_core.cuh"

using namespace mppi_b200;

template <class R, bool FAITHFUL, bool NLE>
static void calc(const double *q, const double *qd, const double *u, double *qdd, double *nle, double *kin) {
    static const RobotModel<R> M = make_robot_model<R>();
    R q_[12], qd_[12], u_[12], qdd_[12], nle_[12];
    for (int i = 0; i < 12; i++) { q_[i] = (R)q[i]; qd_[i] = (R)qd[i]; u_[i] = (R)u[i]; nle_[i] = 0; }
    Kinematics<R> K;
    robot_calculate<R, FAITHFUL, NLE, KIN_MOUNT | KIN_VEL | KIN_MANIP | KIN_LINKS>(M, q_, qd_, u_, qdd_, nle_, K);
    for (int i = 0; i < 12; i++) { qdd[i] = qdd_[i]; nle[i] = nle_[i]; }
    double *k = kin;
    *k++ = K.ee_pos.x; *k++ = K.ee_pos.y; *k++ = K.ee_pos.z;
    *k++ = K.mount_pos.x; *k++ = K.mount_pos.y; *k++ = K.mount_pos.z;
    *k++ = K.ee_lin_vel.x; *k++ = K.ee_lin_vel.y; *k++ = K.ee_lin_vel.z;
    *k++ = K.manip_det;
    for (int l = 0; l < 8; l++) { *k++ = K.link_com[l].x; *k++ = K.link_com[l].y; *k++ = K.link_com[l].z; }
}

// what the FUSED rollout step runs: kinematics / RNEA crossing joints 0..9 by their structure (PLANE) with shared joint
// sines / cosines, then the structure-exploiting solver (EE variant of the lean kernel checked through `ee3`)
template <class R>
static void calc_plane(const double *q, const double *qd, const double *u, double *qdd, double *nle, double *kin, double *ee3) {
    static const RobotModel<R> M = make_robot_model<R>();
    static const FastModel<R> F = make_fast_model<R>();
    R q_[12], qd_[12], u_[12], qdd_[12], nle_[12], cs[12], sn[12];
    for (int i = 0; i < 12; i++) { q_[i] = (R)q[i]; qd_[i] = (R)qd[i]; u_[i] = (R)u[i]; nle_[i] = 0; }
    joint_sincos<R>(F, q_, cs, sn);
    Kinematics<R> K;
    robot_calculate<R, false, true, KIN_MOUNT | KIN_VEL | KIN_MANIP | KIN_LINKS, false, true>(M, q_, qd_, u_, qdd_, nle_, K, cs, sn);
    Vec3<R> ee;
    aba_fused_fast<R, 1, true>(F, q_, cs, sn, u_, qdd_, &ee);
    for (int i = 0; i < 12; i++) { qdd[i] = qdd_[i]; nle[i] = nle_[i]; }
    double *k = kin;
    *k++ = K.ee_pos.x; *k++ = K.ee_pos.y; *k++ = K.ee_pos.z;
    *k++ = K.mount_pos.x; *k++ = K.mount_pos.y; *k++ = K.mount_pos.z;
    *k++ = K.ee_lin_vel.x; *k++ = K.ee_lin_vel.y; *k++ = K.ee_lin_vel.z;
    *k++ = K.manip_det;
    for (int l = 0; l < 8; l++) { *k++ = K.link_com[l].x; *k++ = K.link_com[l].y; *k++ = K.link_com[l].z; }
    ee3[0] = ee.x; ee3[1] = ee.y; ee3[2] = ee.z;
}

// the ROLLED passes (arm joints as one loop body) must reproduce the unrolled structure-crossing passes
template <class R>
static void calc_rolled(const double *q, const double *qd, double *nle, double *kin) {
    static const RobotModel<R> M = make_robot_model<R>();
    static const FastModel<R> F = make_fast_model<R>();
    R q_[12], qd_[12], nle_[12], cs[12], sn[12];
    for (int i = 0; i < 12; i++) { q_[i] = (R)q[i]; qd_[i] = (R)qd[i]; nle_[i] = 0; cs[i] = 1; sn[i] = 0; }
    joint_sincos<R>(F, q_, cs, sn);
    Kinematics<R> K;
    robot_calculate_rolled<R, true, KIN_MOUNT | KIN_VEL | KIN_MANIP | KIN_LINKS>(M, q_, qd_, nle_, K, cs, sn);
    for (int i = 0; i < 12; i++) nle[i] = nle_[i];
    double *k = kin;
    *k++ = K.ee_pos.x; *k++ = K.ee_pos.y; *k++ = K.ee_pos.z;
    *k++ = K.mount_pos.x; *k++ = K.mount_pos.y; *k++ = K.mount_pos.z;
    *k++ = K.ee_lin_vel.x; *k++ = K.ee_lin_vel.y; *k++ = K.ee_lin_vel.z;
    *k++ = K.manip_det;
    for (int l = 0; l < 8; l++) { *k++ = K.link_com[l].x; *k++ = K.link_com[l].y; *k++ = K.link_com[l].z; }
}

extern "C" {
void host_robot_calculate_rolled(int f32, const double *q, const double *qd, double *nle, double *kin34) {
    if (f32) calc_rolled<float>(q, qd, nle, kin34); else calc_rolled<double>(q, qd, nle, kin34);
}
void host_robot_calculate_plane(int f32, const double *q, const double *qd, const double *u, double *qdd, double *nle, double *kin34, double *ee3) {
    if (f32) calc_plane<float>(q, qd, u, qdd, nle, kin34, ee3); else calc_plane<double>(q, qd, u, qdd, nle, kin34, ee3);
}
// mode: bit0 = FAITHFUL, bit1 = NLE requested, bit2 = float
void host_robot_calculate(int mode, const double *q, const double *qd, const double *u, double *qdd, double *nle, double *kin34) {
    switch (mode) {
        case 0: calc<double, false, false>(q, qd, u, qdd, nle, kin34); break;
        case 1: calc<double, true, true>(q, qd, u, qdd, nle, kin34); break;
        case 2: calc<double, false, true>(q, qd, u, qdd, nle, kin34); break;
        case 4: calc<float, false, false>(q, qd, u, qdd, nle, kin34); break;
        case 5: calc<float, true, true>(q, qd, u, qdd, nle, kin34); break;
        case 6: calc<float, false, true>(q, qd, u, qdd, nle, kin34); break;
    }
}
// robot_fast.cuh: qdd = M^-1 u and the end effector position
void host_fast_aba(int f32, const double *q, const double *u, double *qdd, double *ee) {
    if (f32) {
        static const FastModel<float> F = make_fast_model<float>();
        float q_[12], u_[12], c[12], s[12], o[12];
        for (int i = 0; i < 12; i++) { q_[i] = (float)q[i]; u_[i] = (float)u[i]; c[i] = cosf(q_[i]); s[i] = sinf(q_[i]); }
        aba_fused_fast<float>(F, q_, c, s, u_, o);
        Vec3<float> p = ee_position_fast<float>(F, q_, c, s);
        for (int i = 0; i < 12; i++) qdd[i] = o[i];
        ee[0] = p.x; ee[1] = p.y; ee[2] = p.z;
    } else {
        static const FastModel<double> F = make_fast_model<double>();
        double c[12], s[12];
        for (int i = 0; i < 12; i++) { c[i] = cos(q[i]); s[i] = sin(q[i]); }
        aba_fused_fast<double>(F, q, c, s, u, qdd);
        Vec3<double> p = ee_position_fast<double>(F, q, c, s);
        ee[0] = p.x; ee[1] = p.y; ee[2] = p.z;
    }
}
// the unrolled solver (UNROLL = 7: offsets with their structural zeros in the backward pass too), FP64
void host_fast_aba_unrolled(const double *q, const double *u, double *qdd, double *ee) {
    static const FastModel<double> F = make_fast_model<double>();
    double c[12], s[12];
    for (int i = 0; i < 12; i++) { c[i] = cos(q[i]); s[i] = sin(q[i]); }
    Vec3<double> p;
    aba_fused_fast<double, 7, true>(F, q, c, s, u, qdd, &p);
    ee[0] = p.x; ee[1] = p.y; ee[2] = p.z;
}
// the device's FP64 sine / cosine (polynomial core of robot_fast.cuh) evaluated on the host
void host_sincos_poly(int n, const double *a, double *s, double *c) {
    static const FastModel<double> F = make_fast_model<double>();
    for (int i = 0; i < n; i++) sincos_poly(F.trig, a[i], &s[i], &c[i]);
}
void host_sincos_poly_f32(int n, const float *a, float *s, float *c) {
    for (int i = 0; i < n; i++) sincos_poly((const float *)nullptr, a[i], &s[i], &c[i]);
}
// the integer-pipe sign test of the joint-limit term (spatial.cuh)
void host_is_negative(int n, const double *x, int *out64, int *out32) {
    for (int i = 0; i < n; i++) { out64[i] = is_negative(x[i]) ? 1 : 0; out32[i] = is_negative((float)x[i]) ? 1 : 0; }
}
// cost.hpp functors as the kernels evaluate them (rollout_core.cuh): kind 0 quadratic {a,b,c}, 1 left inverse {bound, scale, max},
// 2 right inverse, 3 upper log {bound, scale, offset, max}, 4 lower log; f32 = in single precision
void host_cost_functor(int kind, int f32, double a, double b, double c, double d, const double *values, long count, double *out) {
    for (long i = 0; i < count; i++) {
        if (f32) {
            const float v = (float)values[i];
            switch (kind) {
                case 0: out[i] = quadratic(QuadP<float>{(float)a, (float)b, (float)c}, v); break;
                case 1: out[i] = left_barrier(BarrierP<float>{(float)a, (float)b, (float)c}, v); break;
                case 2: out[i] = right_barrier(BarrierP<float>{(float)a, (float)b, (float)c}, v); break;
                case 3: out[i] = upper_log_barrier(LogBarrierP<float>{(float)a, (float)b, (float)c, (float)d}, v); break;
                default: out[i] = lower_log_barrier(LogBarrierP<float>{(float)a, (float)b, (float)c, (float)d}, v); break;
            }
        } else {
            const double v = values[i];
            switch (kind) {
                case 0: out[i] = quadratic(QuadP<double>{a, b, c}, v); break;
                case 1: out[i] = left_barrier(BarrierP<double>{a, b, c}, v); break;
                case 2: out[i] = right_barrier(BarrierP<double>{a, b, c}, v); break;
                case 3: out[i] = upper_log_barrier(LogBarrierP<double>{a, b, c, d}, v); break;
                default: out[i] = lower_log_barrier(LogBarrierP<double>{a, b, c, d}, v); break;
            }
        }
    }
}
int host_fast_structure_matches() { std::string w; return fast_structure_matches(&w) ? 1 : 0; }
int host_topology_matches() { std::string w; return topology_matches(&w) ? 1 : 0; }
}

// ---- whole rollouts ------------------------------------------------------------------------------
#include "params_convert.h"
#include "rollout_core.cuh"

template <class R, int VAR, bool FAITHFUL, class CP, bool BIG = false>
static void run_rollouts(const CP &cp, const double *x0, const double *U, const double *W, const double *eps, int K, int T, double dt, double discount, double *costs, double *bd) {
    static const RobotModel<R> M = make_robot_model<R>();
    static const FastModel<R> F = make_fast_model<R>();
    static const FastModel<double> F64 = make_fast_model<double>();
    auto P = convert<R>(cp);
    std::vector<R> x(31), u((size_t)12 * T), w, e((size_t)12 * T);
    for (int i = 0; i < 31; i++) x[i] = (R)x0[i];
    for (size_t i = 0; i < u.size(); i++) u[i] = (R)U[i];
    if (W) { w.resize((size_t)6 * T); for (size_t i = 0; i < w.size(); i++) w[i] = (R)W[i]; }
    RolloutInputs<R> in;
    in.x0 = x.data(); in.x0_64 = x0; in.U = u.data(); in.U64 = U; in.W = W ? w.data() : nullptr; in.T = T; in.dt = (R)dt; in.dt64 = dt; in.discount = discount;
    for (int k = 0; k < K; k++) {
        for (size_t i = 0; i < e.size(); i++) e[i] = (R)eps[(size_t)k * 12 * T + i];
        costs[k] = rollout_franka<R, VAR, FAITHFUL, decltype(P), BIG>(M, F, P, in, e.data(), bd, &F64);
    }
}

extern "C" {
// objective: 1 track point, 2 assisted manipulation; flags: bit0 FAITHFUL, bit2 float, bit3 the unrolled build of the lean
// reach-to-pose kernel (k_rollout.cuh BIG: the arm joints' loop index is a constant, so the offsets' structural zeros apply)
void host_rollouts(int objective, int flags, const void *params, const double *x0, const double *U, const double *W, const double *eps, int K, int T,
                   double dt, double discount, double *costs, double *bd7) {
    const bool faithful = flags & 1, f32 = flags & 4;
#define RUN(R, VAR, CP) do { if (faithful) run_rollouts<R, VAR, true>(CP, x0, U, W, eps, K, T, dt, discount, costs, bd7); \
                             else run_rollouts<R, VAR, false>(CP, x0, U, W, eps, K, T, dt, discount, costs, bd7); } while (0)
    if (objective == 1) {
        const auto &cp = *static_cast<const mppi_b200_track_point *>(params);
        int var = variant_for(cp);
        if ((flags & 8) && var == VAR_TP_LEAN && !faithful) {
            if (f32) run_rollouts<float, VAR_TP_LEAN, false, mppi_b200_track_point, true>(cp, x0, U, W, eps, K, T, dt, discount, costs, bd7);
            else run_rollouts<double, VAR_TP_LEAN, false, mppi_b200_track_point, true>(cp, x0, U, W, eps, K, T, dt, discount, costs, bd7);
            return;
        }
        if (f32) { if (var == VAR_TP_LEAN) RUN(float, VAR_TP_LEAN, cp); else RUN(float, VAR_TP_FULL, cp); }
        else { if (var == VAR_TP_LEAN) RUN(double, VAR_TP_LEAN, cp); else RUN(double, VAR_TP_FULL, cp); }
    } else {
        const auto &cp = *static_cast<const mppi_b200_assisted_manipulation *>(params);
        int var = variant_for(cp);
        if (f32) { if (var == VAR_AM) RUN(float, VAR_AM, cp); else RUN(float, VAR_AM_ENERGY, cp); }
        else { if (var == VAR_AM) RUN(double, VAR_AM, cp); else RUN(double, VAR_AM_ENERGY, cp); }
    }
#undef RUN
}
}
