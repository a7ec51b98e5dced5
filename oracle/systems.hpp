// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header).
// The concrete (Dynamics, Cost) pairs of the hot path, restated in FP64:
//   cost functors        — reference src/controller/cost.hpp:10-167
//   energy tank          — src/controller/energy.hpp:5-51
//   PinocchioDynamics    — src/frankaridgeback/pinocchio_dynamics.cpp:142-260 (+ accessors .hpp:138-262)
//   TrackPoint           — src/frankaridgeback/objective/track_point.cpp:10-174
//   AssistedManipulation — src/frankaridgeback/objective/assisted_manipulation.cpp:37-319
//   toy double integrator — NEW (BASELINE.json config 1), no reference counterpart.
// Parameter structs are the plain-C ones of include/mppi_b200.h (interface only).
#pragma once
#include <array>
#include <cstring>

#include "../include/mppi_b200.h"
#include "mppi_oracle.hpp"
#include "robot_oracle.hpp"

namespace oracle {

// ---- cost.hpp ---------------------------------------------------------------------------------
template <class S> inline S quadratic_cost(const mppi_b200_quadratic &c, S value) {  // cost.hpp:25-31
    return S(c.constant_cost) + S(c.linear_cost) * m_fabs(value) + S(c.quadratic_cost) * value * value;
}
template <class S> inline S right_inverse_barrier(const mppi_b200_barrier &b, S value) {  // cost.hpp:57-62
    if (value >= S(b.bound)) return S(b.maximum_cost) + S(b.scale) * m_pow(value - S(b.bound), 2);
    return m_min(S(b.scale) / (S(b.bound) - value), S(b.maximum_cost));
}
template <class S> inline S left_inverse_barrier(const mppi_b200_barrier &b, S value) {  // cost.hpp:88-93
    if (value <= S(b.bound)) return S(b.maximum_cost) + S(b.scale) * m_pow(S(b.bound) - value, 2);
    return m_min(S(b.scale) / (value - S(b.bound)), S(b.maximum_cost));
}
// cost.hpp:105-167 — defined by the reference, used by no objective; kept for coverage.
struct LogBarrier { double bound, scale, offset, maximum_cost = 1e10; };
inline double upper_log_barrier(const LogBarrier &b, double value) {
    if (value >= b.bound) return b.maximum_cost;
    return std::min(b.scale * (-std::log10(-value + b.bound) + b.offset), 0.0);
}
inline double lower_log_barrier(const LogBarrier &b, double value) {
    if (value <= b.bound) return b.maximum_cost;
    return std::min(b.scale * (-std::log10(value - b.bound) + b.offset), 0.0);
}

// ---- energy.hpp -------------------------------------------------------------------------------
template <class S> struct EnergyTank {
    S energy, state;
    void set_energy(S e) { energy = e; state = m_sqrt(S(2.0) * e); }
    void step(S power, S dt) { energy = m_max(S(0.0), energy + power * dt); state = m_sqrt(S(2.0) * energy); }
};

// ---- toy double integrator (NEW) --------------------------------------------------------------
// x = [px, py, vx, vy], u = [ax, ay]; v += u dt; p += v dt (same semi-implicit order as
// pinocchio_dynamics.cpp:245-246).
struct ToyDynamics : Dynamics {
    double x[4] = {0, 0, 0, 0};
    std::unique_ptr<Dynamics> copy() override { return std::make_unique<ToyDynamics>(*this); }
    const double *step(const double *u, double dt) override {
        x[2] += u[0] * dt; x[3] += u[1] * dt;
        x[0] += x[2] * dt; x[1] += x[3] * dt;
        return x;
    }
    void set_state(const double *s, double) override { std::memcpy(x, s, sizeof x); }
    const double *get_state() override { return x; }
    int get_control_dof() override { return 2; }
    int get_state_dof() override { return 4; }
};
struct ToyCost : Cost {
    mppi_b200_toy_objective p;
    explicit ToyCost(const mppi_b200_toy_objective &q) : p(q) {}
    std::unique_ptr<Cost> copy() override { return std::make_unique<ToyCost>(*this); }
    void reset(double) override {}
    double get_cost(const double *s, const double *u, Dynamics *, double) override {
        double ex = s[0] - p.target[0], ey = s[1] - p.target[1];
        return p.position_cost * (ex * ex + ey * ey) + p.velocity_cost * (s[2] * s[2] + s[3] * s[3]) + p.control_cost * (u[0] * u[0] + u[1] * u[1]);
    }
    int get_control_dof() override { return 2; }
    int get_state_dof() override { return 4; }
};

// ---- FrankaRidgeback ---------------------------------------------------------------------------
constexpr int FR_STATE = 31, FR_CONTROL = 12;  // dof.hpp:63,70

// dynamics.hpp:95-117
template <class S> struct EndEffectorState {
    V3<S> position, linear_velocity, angular_velocity, linear_acceleration, angular_acceleration;
    M3<S> orientation;
    S jacobian[6][FR_NJ];
};

// The arithmetic core of PinocchioDynamics, templated so that it can be op-counted.
template <class S> struct RobotCore {
    RobotData<S> data;
    S q[FR_NJ], v[FR_NJ], tau[FR_NJ], acc[FR_NJ];
    EndEffectorState<S> ee;
    EnergyTank<S> tank;
    S state[FR_STATE];
    S power;
    double time = 0.0;
    V3<S> link_com_world[FR_NLINK];

    RobotCore() { for (int i = 0; i < FR_NJ; i++) q[i] = v[i] = tau[i] = acc[i] = S(0.0); for (auto &s : state) s = S(0.0); tank.set_energy(S(0.0)); power = S(0.0); }

    // pinocchio_dynamics.cpp:142-151
    void set_state(const S *x, double t) {
        time = t;
        for (int i = 0; i < FR_STATE; i++) state[i] = x[i];
        for (int i = 0; i < FR_NJ; i++) { q[i] = state[i]; v[i] = state[FR_NJ + i]; }
        tank.set_energy(state[30]);
        calculate();
    }

    // pinocchio_dynamics.cpp:153-224
    void calculate() {
        nonlinear_effects(data, q, v);
        for (int i = 0; i < FR_NJ; i++) tau[i] += data.nle[i];  // "+=": the torque is not cleared by set_state
        aba(data, q, v, tau);
        for (int i = 0; i < FR_NJ; i++) acc[i] = data.ddq[i];
        forward_kinematics2(data, q, v, acc);
        frame_jacobian_world(data);
        for (int r = 0; r < 6; r++) for (int j = 0; j < FR_NJ; j++) ee.jacobian[r][j] = data.J[r][j];
        S yaw = q[2], cy = m_cos(yaw), sy = m_sin(yaw);
        ee.jacobian[0][0] = cy; ee.jacobian[0][1] = -sy; ee.jacobian[0][2] = S(0.0);
        ee.jacobian[1][0] = sy; ee.jacobian[1][1] = cy;  ee.jacobian[1][2] = S(0.0);
        ee.jacobian[2][0] = S(0.0); ee.jacobian[2][1] = S(0.0); ee.jacobian[2][2] = S(1.0);
        Motion<S> sv = act(data.oMi[FR_EE_PARENT], data.v[FR_EE_PARENT]);
        Motion<S> sa = act(data.oMi[FR_EE_PARENT], data.a[FR_EE_PARENT]);
        ee.position = data.oMf_ee.p;
        ee.orientation = data.oMf_ee.R;
        ee.linear_velocity = sv.v; ee.angular_velocity = sv.w;
        ee.linear_acceleration = sa.v; ee.angular_acceleration = sa.w;
        for (int l = 0; l < FR_NLINK; l++) {
            int j = FR_LINK_JOINT[l];
            if (j < 0) { link_com_world[l] = V3<S>(); continue; }
            V3<S> c{S(FR_COM[j][0]), S(FR_COM[j][1]), S(FR_COM[j][2])};
            link_com_world[l] = data.oMi[j].R * c + data.oMi[j].p;
        }
    }

    // pinocchio_dynamics.cpp:226-260
    const S *step(const S *u, S dt) {
        S yaw = q[2], c = m_cos(yaw), s = m_sin(yaw);
        S v0 = c * u[0] - s * u[1], v1 = s * u[0] + c * u[1];
        v[0] = v0; v[1] = v1; v[2] = u[2];
        for (int i = 0; i < FR_NJ; i++) tau[i] = S(0.0);
        for (int i = 0; i < 7; i++) tau[3 + i] = u[3 + i];
        calculate();
        for (int i = 0; i < FR_NJ; i++) v[i] += acc[i] * dt;
        for (int i = 0; i < FR_NJ; i++) q[i] += v[i] * dt;
        power = S(0.0);
        for (int i = 0; i < FR_NJ; i++) power += tau[i] * v[i];
        tank.step(power, dt);
        state[30] = tank.energy;
        for (int i = 0; i < FR_NJ; i++) { state[i] = q[i]; state[FR_NJ + i] = v[i]; }
        time += m_val(dt);
        return state;
    }
};

struct FrankaDynamics : Dynamics {
    RobotCore<double> core;
    const double *wrench = nullptr;  // forecast table T x 6 (or nullptr = no forecast handle)
    double wrench_t0 = 0.0, wrench_dt = 0.01; int wrench_steps = 0;
    std::unique_ptr<Dynamics> copy() override { return std::make_unique<FrankaDynamics>(*this); }
    const double *step(const double *u, double dt) override { return core.step(u, dt); }
    void set_state(const double *s, double t) override { core.set_state(s, t); }
    const double *get_state() override { return core.state; }
    int get_control_dof() override { return FR_CONTROL; }
    int get_state_dof() override { return FR_STATE; }
};

// Shared wrench table: the facade evaluates forecast->get_end_effector_wrench(t0 + k dt) once per
// update (dynamics.hpp:275-278); every clone reads the same table.
struct WrenchTable {
    std::vector<double> w; double t0 = 0.0, dt = 0.01; bool present = false;
    const double *at(double time) const {
        if (!present) return nullptr;
        long k = std::lround((time - t0) / dt);
        if (k < 0) k = 0;
        if ((std::size_t)k * 6 + 6 > w.size()) k = (long)(w.size() / 6) - 1;
        return &w[(std::size_t)k * 6];
    }
};

// joint pairs of track_point.cpp:81-118 / assisted_manipulation.cpp:90-128 (Link enum values)
static const int COLLISION_PAIRS[20][2] = {
    {3, 6}, {3, 7}, {3, 8}, {3, 9}, {3, 10}, {4, 6}, {4, 7}, {4, 8}, {4, 9}, {4, 10},
    {5, 7}, {5, 8}, {5, 9}, {5, 10}, {6, 8}, {6, 9}, {6, 10}, {7, 9}, {7, 10}, {8, 10}};

template <class S> inline V3<S> link_position(const RobotCore<S> &core, int link, int mode) {
    if (mode == MPPI_B200_LINKS_ZERO) return V3<S>();  // pinocchio_dynamics.hpp:189-192
    return core.link_com_world[link];
}

// objective/track_point.cpp:10-174
template <class S> inline S track_point_cost(const mppi_b200_track_point &p, const S *state, const RobotCore<S> &core) {
    static const double lower_limit[12] = {-2.0, -2.0, -6.28, -2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973, 0.5, 0.5};
    static const double upper_limit[12] = {2.0, 2.0, 6.28, 2.8973, 1.7628, 2.8973, 0.0698, 2.8973, 3.7525, 2.8973, 0.5, 0.5};
    V3<S> point{S(p.point[0]), S(p.point[1]), S(p.point[2])};
    S distance = norm(core.ee.position - point);
    S cost = S(100.0) * m_pow(distance, 2);
    if (p.enable_joint_limits) {
        S c = S(0.0);
        for (int i = 0; i < 10; i++) {
            if (state[i] < S(lower_limit[i])) c += S(1000.0) + S(100000.0) * m_pow(S(lower_limit[i]) - state[i], 2);
            if (state[i] > S(upper_limit[i])) c += S(1000.0) + S(100000.0) * m_pow(state[i] - S(upper_limit[i]), 2);
        }
        cost += c;
    }
    if (p.enable_self_collision_avoidance) {
        S c = S(0.0);
        for (auto &pr : COLLISION_PAIRS) {
            S distance = norm(link_position(core, pr[0], p.link_position_mode) - link_position(core, pr[1], p.link_position_mode));
            S radii = S(p.self_collision_radii[pr[0] - 3] + p.self_collision_radii[pr[1] - 3]);
            c += left_inverse_barrier(p.self_collision_limit, radii - distance);  // track_point.cpp:140 (sign flipped vs assisted)
        }
        cost += c;
    }
    if (p.enable_reach_limits) {
        S yaw = core.state[2], cy = m_cos(yaw), sy = m_sin(yaw);
        V3<S> robot = core.data.oMf_mount.p + V3<S>(cy * S(0.3), sy * S(0.3), S(0.15));
        cost += right_inverse_barrier(p.maximum_reach_limit, norm(core.ee.position - robot));
    }
    return cost;
}

struct Breakdown { double joint = 0, self_collision = 0, workspace = 0, energy = 0, velocity = 0, trajectory = 0, manipulability = 0; };

// objective/assisted_manipulation.cpp:37-319
template <class S> inline S assisted_manipulation_cost(const mppi_b200_assisted_manipulation &p, const S *state, const RobotCore<S> &core,
                                                       const double *wrench /* 6 or nullptr */, Breakdown *bd) {
    S cost = S(0.0);
    if (p.enable_joint_limit) {  // :74-88
        S c = S(0.0);
        for (int i = 0; i < FR_NJ; i++) c += left_inverse_barrier(p.lower_joint_limit[i], state[i]) + right_inverse_barrier(p.upper_joint_limit[i], state[i]);
        if (bd) bd->joint += m_val(c);
        cost += c;
    }
    if (p.enable_self_collision_limit) {  // :90-158
        S c = S(0.0);
        for (auto &pr : COLLISION_PAIRS) {
            S distance = norm(link_position(core, pr[0], p.link_position_mode) - link_position(core, pr[1], p.link_position_mode));
            S radii = S(p.self_collision_radii[pr[0] - 3] + p.self_collision_radii[pr[1] - 3]);
            c += left_inverse_barrier(p.self_collision_limit, distance - radii);
        }
        if (bd) bd->self_collision += m_val(c);
        cost += c;
    }
    if (p.enable_workspace_limit) {  // :160-209
        S c = S(0.0);
        V3<S> end_effector = core.ee.position;
        S yaw = core.state[2], cy = m_cos(yaw), sy = m_sin(yaw);
        V3<S> forward{cy, sy, S(0.0)};
        V3<S> robot = core.data.oMf_mount.p + V3<S>(cy * S(0.1), sy * S(0.1), S(0.15));
        V3<S> to_ee = end_effector - robot;
        S projection = dot(to_ee, forward) / dot(forward, forward);
        c += left_inverse_barrier(p.workspace_limit_infront, projection);
        c += right_inverse_barrier(p.workspace_limit_reach, norm(to_ee));
        S n1 = m_sqrt(to_ee.x * to_ee.x + to_ee.y * to_ee.y), n2 = m_sqrt(forward.x * forward.x + forward.y * forward.y);
        S yaw_between = m_acos((to_ee.x * forward.x + to_ee.y * forward.y) / n1 / n2);
        if (!m_isnan(yaw_between)) c += quadratic_cost(p.workspace_cost_yaw, m_fabs(yaw_between));
        c += left_inverse_barrier(p.workspace_limit_above, end_effector.z - robot.z);
        if (bd) bd->workspace += m_val(c);
        cost += c;
    }
    if (p.enable_energy_limit) {  // :211-222
        S e = core.tank.energy;
        S c = left_inverse_barrier(p.energy_limit_below, e) + right_inverse_barrier(p.energy_limit_above, e);
        if (bd) bd->energy += m_val(c);
        cost += c;
    }
    if (p.enable_velocity_cost) {  // :224-235
        S c = S(0.0);
        for (int i = 0; i < FR_NJ; i++) c += S(p.velocity_cost[i].quadratic_cost) * m_pow(m_fabs(state[FR_NJ + i]), 2);
        if (bd) bd->velocity += m_val(c);
        cost += c;
    }
    if (p.enable_trajectory_cost && wrench) {  // :237-290
        S c = S(0.0);
        S mx = S(p.trajectory_target_maximum);
        V3<S> tv{m_max(m_min(S(p.trajectory_target_scale) * S(wrench[0]), mx), -mx),
                 m_max(m_min(S(p.trajectory_target_scale) * S(wrench[1]), mx), -mx),
                 m_max(m_min(S(p.trajectory_target_scale) * S(wrench[2]), mx), -mx)};
        S distance = norm(tv);
        if (distance > S(p.trajectory_position_threshold)) {
            c += quadratic_cost(p.trajectory_position_cost, distance);
            S projection = dot(core.ee.linear_velocity, tv) / dot(tv, tv);
            projection = m_copysign(S(1.0), projection) * norm(tv * projection);
            S target = m_exp(S(p.trajectory_velocity_dropoff) * distance) - S(1.0);
            target = m_min(m_max(target, S(p.trajectory_velocity_minimum)), S(p.trajectory_velocity_maximum));  // std::clamp
            c += quadratic_cost(p.trajectory_velocity_cost, m_fabs(target - projection));
        }
        if (bd) bd->trajectory += m_val(c);
        cost += c;
    }
    if (p.enable_manipulability_cost) {  // :292-319
        S JJt[3][3];
        for (int r = 0; r < 3; r++)
            for (int c2 = 0; c2 < 3; c2++) {
                S s = S(0.0);
                for (int j = 3; j < 10; j++) s += core.ee.jacobian[r][j] * core.ee.jacobian[c2][j];
                JJt[r][c2] = s;
            }
        S det = JJt[0][0] * (JJt[1][1] * JJt[2][2] - JJt[1][2] * JJt[2][1]) - JJt[0][1] * (JJt[1][0] * JJt[2][2] - JJt[1][2] * JJt[2][0]) +
                JJt[0][2] * (JJt[1][0] * JJt[2][1] - JJt[1][1] * JJt[2][0]);
        S volume = m_sqrt(det);
        if (m_isnan(volume)) volume = S(1e-5);
        else volume = m_min(m_max(volume, S(1e-5)), S(1e5));
        S c = quadratic_cost(p.manipulability_cost, S(1.0) / volume);
        if (bd) bd->manipulability += m_val(c);
        cost += c;
    }
    return cost;
}

struct TrackPointCost : Cost {
    mppi_b200_track_point p;
    explicit TrackPointCost(const mppi_b200_track_point &q) : p(q) {}
    std::unique_ptr<Cost> copy() override { return std::make_unique<TrackPointCost>(*this); }
    void reset(double) override {}
    double get_cost(const double *s, const double *, Dynamics *d, double) override {
        return track_point_cost<double>(p, s, static_cast<FrankaDynamics *>(d)->core);
    }
    int get_control_dof() override { return FR_CONTROL; }
    int get_state_dof() override { return FR_STATE; }
};

struct AssistedManipulationCost : Cost {
    mppi_b200_assisted_manipulation p;
    std::shared_ptr<WrenchTable> table;
    Breakdown bd;
    AssistedManipulationCost(const mppi_b200_assisted_manipulation &q, std::shared_ptr<WrenchTable> t) : p(q), table(std::move(t)) {}
    std::unique_ptr<Cost> copy() override { return std::make_unique<AssistedManipulationCost>(*this); }
    void reset(double) override { bd = Breakdown(); }
    double get_cost(const double *s, const double *, Dynamics *d, double time) override {
        return assisted_manipulation_cost<double>(p, s, static_cast<FrankaDynamics *>(d)->core, table ? table->at(time) : nullptr, &bd);
    }
    int get_control_dof() override { return FR_CONTROL; }
    int get_state_dof() override { return FR_STATE; }
};

}  // namespace oracle
