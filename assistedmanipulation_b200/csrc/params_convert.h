// Host-side conversion of the C-ABI objective structs (include/mppi_b200.h) into the parameter
// blocks the kernels read, in the kernel's arithmetic.
#pragma once
#include "../../include/mppi_b200.h"
#include "rollout_core.cuh"

namespace mppi_b200 {

template <class R> inline BarrierP<R> cvt(const mppi_b200_barrier &b) { BarrierP<R> o; o.bound = (R)b.bound; o.scale = (R)b.scale; o.maxc = (R)b.maximum_cost; return o; }
template <class R> inline QuadP<R> cvt(const mppi_b200_quadratic &q) { QuadP<R> o; o.c0 = (R)q.constant_cost; o.c1 = (R)q.linear_cost; o.c2 = (R)q.quadratic_cost; return o; }

template <class R> inline void pair_radii(const double *radii8, R *out20) {
    for (int i = 0; i < 20; i++) out20[i] = (R)(radii8[MPPI_PAIR_A(i)] + radii8[MPPI_PAIR_B(i)]);
}

template <class R> inline ToyP<R> convert(const mppi_b200_toy_objective &p) {
    ToyP<R> o; o.target[0] = (R)p.target[0]; o.target[1] = (R)p.target[1]; o.qp = (R)p.position_cost; o.qv = (R)p.velocity_cost; o.qu = (R)p.control_cost; return o;
}
template <class R> inline TrackPointP<R> convert(const mppi_b200_track_point &p) {
    TrackPointP<R> o;
    for (int i = 0; i < 3; i++) o.point[i] = (R)p.point[i];
    static const double lo[10] = {-2.0, -2.0, -6.28, -2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};   // track_point.cpp:48-65
    static const double hi[10] = {2.0, 2.0, 6.28, 2.8973, 1.7628, 2.8973, 0.0698, 2.8973, 3.7525, 2.8973};
    for (int i = 0; i < 10; i++) { o.lim_lo[i] = (R)lo[i]; o.lim_hi[i] = (R)hi[i]; o.lim_lo64[i] = lo[i]; o.lim_hi64[i] = hi[i]; }
    o.joint_limits = p.enable_joint_limits; o.self_collision = p.enable_self_collision_avoidance; o.reach = p.enable_reach_limits;
    o.link_mode = p.link_position_mode;
    o.collision_limit = cvt<R>(p.self_collision_limit);
    pair_radii<R>(p.self_collision_radii, o.radii);
    o.reach_limit = cvt<R>(p.maximum_reach_limit);
    return o;
}
template <class R> inline AssistedP<R> convert(const mppi_b200_assisted_manipulation &p) {
    AssistedP<R> o;
    o.joint_limit = p.enable_joint_limit; o.self_collision = p.enable_self_collision_limit; o.workspace = p.enable_workspace_limit;
    o.energy = p.enable_energy_limit; o.velocity = p.enable_velocity_cost; o.trajectory = p.enable_trajectory_cost;
    o.manipulability = p.enable_manipulability_cost; o.link_mode = p.link_position_mode;
    for (int i = 0; i < 12; i++) { o.lower[i] = cvt<R>(p.lower_joint_limit[i]); o.upper[i] = cvt<R>(p.upper_joint_limit[i]); o.vel_quad[i] = (R)p.velocity_cost[i].quadratic_cost; }
    for (int i = 0; i < 12; i++) { o.lower64[i] = p.lower_joint_limit[i].bound; o.upper64[i] = p.upper_joint_limit[i].bound; }
    o.energy_below64 = p.energy_limit_below.bound; o.energy_above64 = p.energy_limit_above.bound;
    o.collision_limit = cvt<R>(p.self_collision_limit);
    pair_radii<R>(p.self_collision_radii, o.radii);
    o.ws_above = cvt<R>(p.workspace_limit_above); o.ws_infront = cvt<R>(p.workspace_limit_infront); o.ws_reach = cvt<R>(p.workspace_limit_reach);
    o.ws_yaw = cvt<R>(p.workspace_cost_yaw);
    o.energy_below = cvt<R>(p.energy_limit_below); o.energy_above = cvt<R>(p.energy_limit_above);
    o.traj_scale = (R)p.trajectory_target_scale; o.traj_max = (R)p.trajectory_target_maximum; o.traj_threshold = (R)p.trajectory_position_threshold;
    o.traj_vmin = (R)p.trajectory_velocity_minimum; o.traj_vmax = (R)p.trajectory_velocity_maximum; o.traj_dropoff = (R)p.trajectory_velocity_dropoff;
    o.traj_position = cvt<R>(p.trajectory_position_cost); o.traj_velocity = cvt<R>(p.trajectory_velocity_cost); o.manip = cvt<R>(p.manipulability_cost);
    return o;
}

// which kernel variant serves an objective configuration
inline int variant_for(const mppi_b200_track_point &p) { return (p.enable_self_collision_avoidance || p.enable_reach_limits) ? VAR_TP_FULL : VAR_TP_LEAN; }
inline int variant_for(const mppi_b200_assisted_manipulation &p) { return p.enable_energy_limit ? VAR_AM_ENERGY : VAR_AM; }

}  // namespace mppi_b200
