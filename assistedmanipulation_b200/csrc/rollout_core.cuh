// One MPPI rollout, start to finish, for one thread: the T-step recurrence of
// mppi::Trajectory::rollout(Rollout*, Dynamics*, Cost*) (reference src/controller/mppi.cpp:309-342)
// with the objective (objective/track_point.cpp:10-174, objective/assisted_manipulation.cpp:37-319,
// cost.hpp functors, energy.hpp tank) and the dynamics step (pinocchio_dynamics.cpp:226-260) fused
// into straight-line code. Joint state, tank energy and the one-step-stale kinematics the costs
// read stay in registers for the whole horizon; the only memory traffic is
// the rollout's own noise row and the shared control / wrench tables.
// Host/device: tests compile this for the CPU to check it against the oracle.
#pragma once
#include "robot.cuh"
#include "robot_fast.cuh"

// FP32 fast mode: 0 = everything single precision; 1 = the state (q, v, tank energy) carried and integrated in FP64 with
// FP64 barrier sides; 2 = additionally the whole state path in FP64 for the objectives with barrier steps on kinematic
// quantities (see MIXED_SOLVER in rollout_franka). A/B builds set it on the command line.
#ifndef MPPI_MIXED_STATE
#define MPPI_MIXED_STATE 2
#endif
// 1 = the kernels that evaluate kinematics + RNEA every step run those passes, the collision pairs and the joint barriers
// as loops (robot.cuh robot_calculate_rolled) — a third of the code for kernels bound by instruction fetch. MEASURED AND
// REJECTED (B200, config 3): rollout stage 988 us unrolled, 2245 us rolled (one warp 2018; loop-body FP64 solver on top
// 2229): the per-joint results move from registers to thread-local arrays (3.2 KB frames x 224 threads per SM do not
// fit L1) and each loop iteration is one dependent chain the scheduler cannot interleave with its neighbours. Kept, off,
// with its host test (identical values), as the record of that measurement.
#ifndef MPPI_ROLLED
#define MPPI_ROLLED 0
#endif
#ifndef MPPI_MIXED_SOLVER_UNROLL
#define MPPI_MIXED_SOLVER_UNROLL 7
#endif

namespace mppi_b200 {

template <class R> struct BarrierP { R bound, scale, maxc; };
template <class R> struct QuadP { R c0, c1, c2; };

// cost.hpp:25-31
template <class R> MPPI_HD R quadratic(const QuadP<R> &c, R v) { return c.c0 + c.c1 * fabs_(v) + c.c2 * v * v; }
// cost.hpp:57-62 and :88-93. Same value in every case as the reference's two-branch form, written as selects: a warp
// whose rollouts sit on both sides of a bound does not run both paths one after the other, and — what matters more with
// ~50 barrier terms per step — there is no control flow at all, so the terms are one basic block that the scheduler
// overlaps (as branches each term was its own chain of compare, branch, reciprocal, compare, branch: ~100 cycles, 46 % of
// the assisted-manipulation kernel's issue time in the static model).
//   FP32 (the fast mode): scale / x as scale * rcp(x), one MUFU and one multiply (2 ulp; the compiler's approximate
//   division adds four range-scaling multiplies and two comparisons). x is kept at or above the smallest normal number,
//   so the reciprocal is finite and a ZERO scale needs no special case: 0 * rcp(x) is the reference's 0 / x = 0 for
//   every x > 0, and NaN for a NaN state (the comparison-select keeps NaN, unlike fmaxf).
//   FP64: IEEE division, skipped by a uniform branch when the scale is zero (the parameters are constants) — 0 / x is 0.
MPPI_HD float barrier_div(float scale, float x) {
    const float xs = (x < 1.17549435e-38f) ? 1.17549435e-38f : x;
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xs));
    return scale * r;
#else
    return scale / xs;
#endif
}
MPPI_HD double barrier_div(double scale, double x) { return scale / x; }
template <class R> MPPI_HD R barrier_inside(const BarrierP<R> &b, R v, R x) {   // x: distance to the bound, positive inside
    if (sizeof(R) == 8 && b.scale == R(0)) return (v != v) ? v : std_min(R(0), b.maxc);
    return std_min(barrier_div(b.scale, x), b.maxc);
}
template <class R> MPPI_HD R right_barrier(const BarrierP<R> &b, R v) {
    const R d = v - b.bound;
    const R outside = b.maxc + b.scale * (d * d);
    const R inside = barrier_inside(b, v, b.bound - v);
    return (v >= b.bound) ? outside : inside;
}
template <class R> MPPI_HD R left_barrier(const BarrierP<R> &b, R v) {
    const R d = b.bound - v;
    const R outside = b.maxc + b.scale * (d * d);
    const R inside = barrier_inside(b, v, v - b.bound);
    return (v <= b.bound) ? outside : inside;
}

// cost.hpp:105-167 — UpperLogarithmicBarrierFunction / LowerLogarithmicBarrierFunction. Defined by the reference and used
// by none of its objectives; provided for objectives built on this engine, checked value for value
// against the reference's own functors (tests/test_device_math_host.py). log10_ : spatial.cuh.
template <class R> struct LogBarrierP { R bound, scale, offset, maxc; };
template <class R> MPPI_HD R upper_log_barrier(const LogBarrierP<R> &b, R v) {
    if (v >= b.bound) return b.maxc;
    return std_min(b.scale * (-log10_(-v + b.bound) + b.offset), R(0));
}
template <class R> MPPI_HD R lower_log_barrier(const LogBarrierP<R> &b, R v) {
    if (v <= b.bound) return b.maxc;
    return std_min(b.scale * (-log10_(v - b.bound) + b.offset), R(0));
}

// FP32 fast mode, barriers on a STATE variable (joint positions, tank energy). The rollout carries those in FP64 (see
// rollout_franka): which side of the bound the state is on is decided on the FP64 value against the FP64 bound — the
// reference's 1e10 steps (cost.hpp:59-61,90-92) are then taken by exactly the rollouts that take them in FP64 unless the
// state itself differs — while the magnitudes (distance squared, scale / distance) stay single precision.
// d = v - bound (right) or bound - v (left): not negative = at or beyond the bound. The sign is read on the integer pipe.
template <class R> MPPI_HD R barrier_signed(const BarrierP<R> &b, double d) {
    const R df = (R)d;
    const R outside = b.maxc + b.scale * (df * df);
    const R inside = std_min(barrier_div(b.scale, -df), b.maxc);
    return is_negative(d) ? inside : outside;
}

// ---- objective parameter blocks in kernel arithmetic -----------------------------------------
template <class R> struct ToyP { R target[2]; R qp, qv, qu; };

template <class R> struct TrackPointP {
    R point[3];
    R lim_lo[10], lim_hi[10];   // track_point.cpp:48-65 (hard-coded there); here in the parameter block so they are constant-bank operands
    double lim_lo64[10], lim_hi64[10];   // the same limits unrounded: the FP32 fast mode tests its FP64 joint positions against these
    int joint_limits, self_collision, reach, link_mode;
    BarrierP<R> collision_limit;
    R radii[20];  // sum of the two sphere radii per checked pair
    BarrierP<R> reach_limit;
};

template <class R> struct AssistedP {
    int joint_limit, self_collision, workspace, energy, velocity, trajectory, manipulability, link_mode;
    BarrierP<R> lower[12], upper[12];
    BarrierP<R> collision_limit;
    R radii[20];
    BarrierP<R> ws_above, ws_infront, ws_reach;
    QuadP<R> ws_yaw;
    BarrierP<R> energy_below, energy_above;
    R vel_quad[12];
    R traj_scale, traj_max, traj_threshold, traj_vmin, traj_vmax, traj_dropoff;
    QuadP<R> traj_position, traj_velocity, manip;
    double lower64[12], upper64[12], energy_below64, energy_above64;   // unrounded bounds for the FP32 fast mode's FP64 state (barrier_signed)
};

// pairs of Link enum values minus 3 (PIVOT = 0 ... PANDA_LINK7 = 7): track_point.cpp:81-118
#define MPPI_PAIR_A(i) ((i) < 5 ? 0 : (i) < 10 ? 1 : (i) < 14 ? 2 : (i) < 17 ? 3 : (i) < 19 ? 4 : 5)
#define MPPI_PAIR_B(i) ((i) < 5 ? 3 + (i) : (i) < 10 ? 3 + (i) - 5 : (i) < 14 ? 4 + (i) - 10 : (i) < 17 ? 5 + (i) - 14 : (i) < 19 ? 6 + (i) - 17 : 7)

// The link-mode test sits OUTSIDE the pair loop: tested per pair (a uniform branch) every pair was its own basic block
// and the twenty square roots / reciprocals ran one after the other (~48 cycles each in the static model, a sixth of the
// assisted-manipulation step); as one block the scheduler overlaps them. Same operations, same order of additions.
template <class R, bool FLIP> MPPI_HD R self_collision_cost(const BarrierP<R> &lim, const R *radii, int link_mode, const Kinematics<R> &K) {
#if MPPI_ROLLED
    // one loop body for the twenty pairs (same operations, same order of additions; see MPPI_ROLLED)
    R total = R(0);
#pragma unroll 1
    for (int i = 0; i < 20; i++) {
        R dist = R(0);
        if (link_mode != 0) {
            const Vec3<R> d = K.link_com[MPPI_PAIR_A(i)] - K.link_com[MPPI_PAIR_B(i)];
            dist = sqrt_(dot(d, d));
        }
        total += left_barrier(lim, FLIP ? radii[i] - dist : dist - radii[i]);
    }
    return total;
#endif
    R distance[20];
    if (link_mode != 0) {
#pragma unroll
        for (int i = 0; i < 20; i++) {
            const Vec3<R> d = K.link_com[MPPI_PAIR_A(i)] - K.link_com[MPPI_PAIR_B(i)];
            distance[i] = sqrt_(dot(d, d));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 20; i++) distance[i] = R(0);
    }
    R cost = R(0);
#pragma unroll
    for (int i = 0; i < 20; i++) {
        // track_point.cpp:140 uses radii - distance, assisted_manipulation.cpp:149 distance - radii
        cost += left_barrier(lim, FLIP ? radii[i] - distance[i] : distance[i] - radii[i]);
    }
    return cost;
}

enum Variant { VAR_TOY = 0, VAR_TP_LEAN = 1, VAR_TP_FULL = 2, VAR_AM = 3, VAR_AM_ENERGY = 4 };

template <int VAR> struct VariantTraits {
    static constexpr int kin = VAR == VAR_TP_LEAN ? 0 : (VAR == VAR_TP_FULL ? (KIN_MOUNT | KIN_LINKS) : (KIN_MOUNT | KIN_VEL | KIN_MANIP | KIN_LINKS));
    static constexpr bool power = VAR == VAR_AM_ENERGY;
};

// kinematics only (PinocchioDynamics::set_state -> calculate(), the values the first cost evaluation reads)
template <class R, int FLAGS> MPPI_HD void robot_kinematics(const RobotModel<R> &M, const R *q, const R *qd, Kinematics<R> &K) {
    Scratch<R> S;
    Mot<R> agf[1];
    pass1<R, 0, (FLAGS & KIN_VEL) != 0, false, false>(M, q, qd, S, agf);
    Xf<R> oM;
    R Jl[21];
    world_chain<R, 0, FLAGS>(M, S, oM, K, Jl);
    if (FLAGS & KIN_MANIP) {
        R g[6];
        int n = 0;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = r; c < 3; c++) {
                R s = R(0);
#pragma unroll
                for (int j = 0; j < 7; j++) s += Jl[r * 7 + j] * Jl[c * 7 + j];
                g[n++] = s;
            }
        K.manip_det = g[0] * (g[3] * g[5] - g[4] * g[4]) - g[1] * (g[1] * g[5] - g[4] * g[2]) + g[2] * (g[1] * g[4] - g[3] * g[2]);
    }
}

// objective/track_point.cpp:10-79,120-174
// LEAN: the engine picked the variant without self-collision and reach terms, so they are compiled out
// yaw: cos / sin of q[2] when the caller already has them (FUSED mode shares the step's joint sines / cosines), else null
// q64: the FP64 joint positions of the FP32 fast mode (null in FP64 builds, where q is already that)
template <class R, bool LEAN = false> MPPI_HD R track_point_cost(const TrackPointP<R> &P, const R *q, const Kinematics<R> &K, const R *yaw = nullptr, const double *q64 = nullptr) {
    const Vec3<R> e = K.ee_pos - v3<R>(P.point[0], P.point[1], P.point[2]);
    // 100 * |e|^2: the reference squares the norm it took the root of (track_point.cpp:38-41); the root is skipped here
    // (one rounding less, 1e-16 relative)
    R cost = R(100) * dot(e, e);
    if (P.joint_limits) {
        // track_point.cpp:48-65: 1000 + 1e5 * excess^2 for every joint outside [lo, hi]. With the default noise most
        // rollouts are outside some limit for most of the horizon (63 % of the rollout-steps of BASELINE config 2), so
        // there is no fast path to branch to; the term is written for a low FP64 instruction count instead: at most one
        // side is violated, so ONE penalty is formed per joint, of the negative one of (q - lo, hi - q) — its square is
        // the square of the reference's (lo - q) or (q - hi) — and both selections read the sign bit on the integer
        // pipe (an FP64 comparison occupies the FP64 pipe and delivers its predicate ~13 cycles later). Same values and
        // order of additions for every state (a NaN joint is "not negative" on either side, as in the reference's
        // comparisons; the end effector term then makes the rollout NaN one step later).
        R c[2] = {R(0), R(0)};   // two partial sums: half the dependent additions (the sum differs from the sequential one by rounding only)
#pragma unroll
        for (int i = 0; i < 10; i++) {
            if constexpr (sizeof(R) == 4) {
                if (q64) {   // sides decided on the FP64 state against the unrounded limits, magnitudes in single precision
                    const double below = q64[i] - P.lim_lo64[i], above = P.lim_hi64[i] - q64[i];
                    const double md = is_negative(below) ? below : above;
                    const R m = (R)md;
                    const R penalty = R(1000) + R(100000) * (m * m);
                    c[i & 1] += is_negative(md) ? penalty : R(0);
                    continue;
                }
            }
            const R below = q[i] - P.lim_lo[i], above = P.lim_hi[i] - q[i];
            const R m = is_negative(below) ? below : above;
            const R penalty = R(1000) + R(100000) * (m * m);
            c[i & 1] += is_negative(m) ? penalty : R(0);
        }
        cost += c[0] + c[1];
    }
    if constexpr (LEAN) return cost;
    if (P.self_collision) cost += self_collision_cost<R, true>(P.collision_limit, P.radii, P.link_mode, K);
    if (P.reach) {
        R sy, cy;
        if (yaw) { cy = yaw[0]; sy = yaw[1]; } else sincos_(q[2], &sy, &cy);
        const Vec3<R> robot = K.mount_pos + v3<R>(cy * R(0.3), sy * R(0.3), R(0.15));
        const Vec3<R> d = K.ee_pos - robot;
        cost += right_barrier(P.reach_limit, sqrt_(dot(d, d)));
    }
    return cost;
}

// objective/assisted_manipulation.cpp:37-319; bd (7 doubles) accumulates the per-term totals the
// reference's logger reads after Trajectory::filter() (logging/assisted_manipulation.cpp:58-103).
// q64 / energy64: the FP64 state of the FP32 fast mode (null in FP64 builds)
template <class R> MPPI_HD R assisted_cost(const AssistedP<R> &P, const R *q, const R *qd, R energy, const Kinematics<R> &K, const R *wrench, double *bd, const R *yaw = nullptr,
                                           const double *q64 = nullptr, double energy64 = 0.0) {
    R cost = R(0);
    if (P.joint_limit) {
        R c = R(0);
        if (sizeof(R) == 4 && q64) {
#if MPPI_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int i = 0; i < NJ; i++) c += barrier_signed(P.lower[i], P.lower64[i] - q64[i]) + barrier_signed(P.upper[i], q64[i] - P.upper64[i]);
        } else {
#if MPPI_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int i = 0; i < NJ; i++) c += left_barrier(P.lower[i], q[i]) + right_barrier(P.upper[i], q[i]);
        }
        if (bd) bd[0] += (double)c;
        cost += c;
    }
    if (P.self_collision) {
        const R c = self_collision_cost<R, false>(P.collision_limit, P.radii, P.link_mode, K);
        if (bd) bd[1] += (double)c;
        cost += c;
    }
    if (P.workspace) {
        R c = R(0);
        R sy, cy;
        if (yaw) { cy = yaw[0]; sy = yaw[1]; } else sincos_(q[2], &sy, &cy);
        const Vec3<R> fwd = v3<R>(cy, sy, R(0));
        const Vec3<R> robot = K.mount_pos + v3<R>(cy * R(0.1), sy * R(0.1), R(0.15));
        const Vec3<R> d = K.ee_pos - robot;
        c += left_barrier(P.ws_infront, dot(d, fwd) / dot(fwd, fwd));
        c += right_barrier(P.ws_reach, sqrt_(dot(d, d)));
        const R n1 = sqrt_(d.x * d.x + d.y * d.y), n2 = sqrt_(fwd.x * fwd.x + fwd.y * fwd.y);
        const R yaw = acos_((d.x * fwd.x + d.y * fwd.y) / n1 / n2);
        if (!(yaw != yaw)) c += quadratic(P.ws_yaw, fabs_(yaw));
        c += left_barrier(P.ws_above, K.ee_pos.z - robot.z);
        if (bd) bd[2] += (double)c;
        cost += c;
    }
    if (P.energy) {
        const R c = (sizeof(R) == 4 && q64) ? barrier_signed(P.energy_below, P.energy_below64 - energy64) + barrier_signed(P.energy_above, energy64 - P.energy_above64)
                                            : left_barrier(P.energy_below, energy) + right_barrier(P.energy_above, energy);
        if (bd) bd[3] += (double)c;
        cost += c;
    }
    if (P.velocity) {
        R c = R(0);
#pragma unroll
        for (int i = 0; i < NJ; i++) { const R a = fabs_(qd[i]); c += P.vel_quad[i] * (a * a); }
        if (bd) bd[4] += (double)c;
        cost += c;
    }
    if (P.trajectory && wrench) {
        R c = R(0);
        const Vec3<R> t = v3<R>(std_max(std_min(P.traj_scale * wrench[0], P.traj_max), -P.traj_max),
                                std_max(std_min(P.traj_scale * wrench[1], P.traj_max), -P.traj_max),
                                std_max(std_min(P.traj_scale * wrench[2], P.traj_max), -P.traj_max));
        const R distance = sqrt_(dot(t, t));
        if (distance > P.traj_threshold) {
            c += quadratic(P.traj_position, distance);
            R proj = dot(K.ee_lin_vel, t) / dot(t, t);
            const Vec3<R> tp = t * proj;
            proj = copysign_(R(1), proj) * sqrt_(dot(tp, tp));
            const R target = std_clamp(exp_(P.traj_dropoff * distance) - R(1), P.traj_vmin, P.traj_vmax);
            c += quadratic(P.traj_velocity, fabs_(target - proj));
        }
        if (bd) bd[5] += (double)c;
        cost += c;
    }
    if (P.manipulability) {
        R vol = sqrt_(K.manip_det);
        if (vol != vol) vol = R(1e-5);
        else vol = std_clamp(vol, R(1e-5), R(1e5));
        const R c = quadratic(P.manip, R(1) / vol);
        if (bd) bd[6] += (double)c;
        cost += c;
    }
    return cost;
}

// ---- noise chase (small rollout sets; k_rollout.cuh) ---------------------------------------------------------------------
// When the rollout grid leaves most of every SM empty (K = 4096: one warp per SM), the sampling kernel's work moves INTO the
// rollout blocks: seven more warps per block draw the noise of the block's own 32 rollouts, chunk by chunk of CHASE_STEPS
// steps, while warp 0 integrates — the rollout starts ~1 us after the launch instead of behind a 10 us sampling kernel, and
// nothing crosses a block. A chunk is complete when the block's shared-memory counter of its columns reaches
// CHASE_COLUMNS; the rollout warp looks at the counter in front of its first load of the chunk. In a launch without
// sampling warps the counters are preset, so the wait needs no test of its own.
constexpr int CHASE_STEPS = 4;
constexpr int CHASE_COLUMNS = 32 * CHASE_STEPS;   // columns of one chunk: 32 rollouts x CHASE_STEPS steps
constexpr int CHASE_MAX_CHUNKS = 256;             // horizons up to 1024 steps
#if defined(__CUDACC__)
__shared__ unsigned s_chase_columns[CHASE_MAX_CHUNKS];   // written by the kernel before its first barrier
// GUARD: give up (trap) after ~1 s — for the wait in front of the step loop. The waits inside the loop are as few
// instructions as possible (the unrolled FP64 step body is 31.1 KB of a 32 KB instruction cache level: 50 more instructions
// in it cost 8 % of the kernel); once chunk 0 has arrived the sampling warps are running, and they wait for nothing.
template <bool GUARD> __device__ __forceinline__ void noise_chase_wait(int chunk) {
    const unsigned address = (unsigned)__cvta_generic_to_shared(s_chase_columns + chunk);
    unsigned seen;
    for (unsigned spins = 0;; spins++) {
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(address) : "memory");
        if (seen == (unsigned)CHASE_COLUMNS) break;
        if (GUARD && spins > (1u << 24)) __trap();
    }
}
// the values must be in registers here (keeps the compiler from sinking their computation below the loads that follow,
// which would cost a copy of the previous noise column per step)
__device__ __forceinline__ void pin_values(double *v) { asm volatile("" : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]), "+d"(v[4]), "+d"(v[5]), "+d"(v[6]), "+d"(v[7]), "+d"(v[8]), "+d"(v[9]), "+d"(v[10]), "+d"(v[11])); }
__device__ __forceinline__ void pin_values(float *v) { asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11])); }
#endif

// one step of a rollout's noise row: 12 values, 16-byte aligned (rows are multiples of 16 bytes). COHERENT: ordinary loads —
// the kernels whose blocks may draw the noise themselves (above) read rows written by another warp during the launch.
template <bool COHERENT = false> MPPI_HD void load_eps(const double *p, double *o) {
#if defined(__CUDA_ARCH__)
    const double2 *v = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int i = 0; i < 6; i++) { const double2 t = COHERENT ? v[i] : __ldg(v + i); o[2 * i] = t.x; o[2 * i + 1] = t.y; }
#else
    for (int i = 0; i < 12; i++) o[i] = p[i];
#endif
}
template <bool COHERENT = false> MPPI_HD void load_eps(const float *p, float *o) {
#if defined(__CUDA_ARCH__)
    const float4 *v = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int i = 0; i < 3; i++) { const float4 t = COHERENT ? v[i] : __ldg(v + i); o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w; }
#else
    for (int i = 0; i < 12; i++) o[i] = p[i];
#endif
}

MPPI_HD double discount_pow(double g, int step) { return g == 1.0 ? 1.0 : pow(g, (double)step); }

// Everything one rollout of the Franka+Ridgeback system needs that does not depend on the sample.
template <class R> struct RolloutInputs {
    const R *x0;      // 31: q, qd, wrench, tank energy (state.hpp:113-260)
    const double *x0_64 = nullptr;   // the same state unrounded (FP32 fast mode)
    const R *U;       // nu x T column-major: m_optimal_control_shifted
    const double *U64 = nullptr;     // the same sequence unrounded (FP32 fast mode with the FP64 state path)
    const R *W;       // T x 6 forecast wrench table, or nullptr (no forecast handle)
    int T;
    R dt;
    double dt64 = 0.0;   // the time step unrounded (FP32 fast mode: its FP64 state integrates with this one)
    double discount;
    const double *discount_table = nullptr;   // pow(discount, step) per step, staged by the kernel (keeps pow() out of the step loop)
};

// VAR selects objective + which kinematics are alive; FAITHFUL selects the dynamics evaluation.
// eps: this rollout's noise, [t][d]. Returns the rollout cost (NaN = failed rollout, mppi.cpp:331-334).
// BIG: the build for rollout sets that fill the machine (see k_rollout.cuh)
// F64: the solver's model in FP64 (FP32 fast mode of the objectives with barrier steps: see MIXED_SOLVER below); null otherwise
template <class R, int VAR, bool FAITHFUL, class ParamsT, bool BIG = false, bool CHASE = false>   // CHASE: see "noise chase" above
MPPI_HD double rollout_franka(const RobotModel<R> &M, const FastModel<R> &F, const ParamsT &P, const RolloutInputs<R> &in, const R *eps, double *bd, const FastModel<double> *F64 = nullptr) {
    constexpr int KF = VariantTraits<VAR>::kin;
    constexpr bool POWER = VariantTraits<VAR>::power;
    constexpr bool LEAN = !FAITHFUL && KF == 0 && !POWER;  // only the end effector position is read by the objective
    // FP32 fast mode: the STATE (joint positions, velocities, tank energy) is carried and integrated in FP64, the dynamics
    // and the objective read single-precision copies of it. Rounding the state to FP32 after every step moved rollouts
    // across the reference's 1e10 barrier steps (cost.hpp:59-61,90-92) — one such "flip" changes the published control
    // sequence by ~1e-4 of its maximum, the whole tolerance of the fast mode. The accelerations still come from the FP32
    // solver; what is removed is the 6e-8 relative rounding of q and v per step, of the bounds and of the time step.
    // (The lean reach-to-pose kernel stays all single precision: its only steps are the 1000-unit joint-limit penalties of
    // track_point.cpp:48-65, a rollout on the other side of one moves the control sequence by ~1e-6, and the FP64 state
    // cost it 7 % at config 4.)
    constexpr bool MIXED = sizeof(R) == 4 && MPPI_MIXED_STATE && !LEAN;
    // Objectives with 1e10 barrier steps on kinematic quantities (assisted manipulation, full reach-to-pose): the whole
    // STATE PATH — control + noise, base velocity, joint sines / cosines, the solver qdd = M(q)^-1 tau, the integration —
    // runs in FP64, so the trajectory is the FP64 kernel's; kinematics, RNEA (tank power) and the objective stay FP32.
    // With the FP32 solver the trajectory drifts by ~1e-6 over the horizon and ~0.07 % of the rollouts end up on the
    // other side of a barrier step than their FP64 twin (measured, config 3: 60 % of those at the collision spheres,
    // 24 % at joint limits, 16 % at the workspace planes), each worth ~1e-4 of the published control sequence.
    constexpr bool MIXED_SOLVER = MIXED && !FAITHFUL && !LEAN && MPPI_MIXED_STATE >= 2;
    R q[NJ], qd[NJ];
    double q64[MIXED ? NJ : 1], qd64[MIXED ? NJ : 1], energy64 = 0.0;
#pragma unroll
    for (int i = 0; i < NJ; i++) { q[i] = in.x0[i]; qd[i] = in.x0[NJ + i]; }
    R energy = in.x0[30];
    if constexpr (MIXED) {
        const double *x64 = in.x0_64;
#pragma unroll
        for (int i = 0; i < NJ; i++) { q64[i] = x64[i]; qd64[i] = x64[NJ + i]; }
        energy64 = x64[30];
    }
    const double *q64p = MIXED ? q64 : nullptr;
    Kinematics<R> K;
    R cs[NJ], sn[NJ];
    // FUSED: one set of joint sines / cosines per state, evaluated when the state is formed (here and after every
    // integration) and read by the objective's yaw terms, the kinematics and the solver of the next step
    double cs64[MIXED_SOLVER ? NJ : 1], sn64[MIXED_SOLVER ? NJ : 1];
    if constexpr (MIXED_SOLVER) {
        joint_sincos<double>(*F64, q64, cs64, sn64);
#pragma unroll
        for (int i = 2; i < 10; i++) { cs[i] = (R)cs64[i]; sn[i] = (R)sn64[i]; }
    } else if constexpr (!FAITHFUL) joint_sincos<R>(F, q, cs, sn);
    if constexpr (LEAN) {
        K.ee_pos = ee_position_fast<R>(F, q, cs, sn);
    } else {
        robot_kinematics<R, KF>(M, q, qd, K);
    }
    double total = 0.0;
    R e_next[NJ];
#if defined(__CUDA_ARCH__)
    if constexpr (CHASE) noise_chase_wait<true>(0);
#endif
    load_eps<CHASE>(eps, e_next);
    for (int step = 0; step < in.T; ++step) {
        R u[NJ];
        double u64[MIXED_SOLVER ? NJ : 1];
        if constexpr (MIXED_SOLVER) {
#pragma unroll
            for (int d = 0; d < NJ; d++) { u64[d] = in.U64[step * NJ + d] + (double)e_next[d]; u[d] = (R)u64[d]; }
        } else {
#pragma unroll
            for (int d = 0; d < NJ; d++) u[d] = in.U[step * NJ + d] + e_next[d];
        }
#if defined(__CUDA_ARCH__)
        if constexpr (CHASE) pin_values(u);
#endif
        if (step + 1 < in.T) load_eps<CHASE>(eps + (step + 1) * NJ, e_next);  // next step's noise is in flight during this step
        R c;
        R yaw[2] = {cs[2], sn[2]};
        const R *yawp = FAITHFUL ? nullptr : yaw;
        if constexpr (VAR == VAR_TP_LEAN) c = track_point_cost<R, true>(P, q, K, nullptr, q64p);
        else if constexpr (VAR == VAR_TP_FULL) c = track_point_cost<R>(P, q, K, yawp, q64p);
        else c = assisted_cost<R>(P, q, qd, energy, K, in.W ? in.W + step * 6 : nullptr, bd, yawp, q64p, energy64);
        const double sc = (in.discount_table ? in.discount_table[step] : discount_pow(in.discount, step)) * (double)c;
        // A NaN stage cost ends the reference's rollout with a NaN total (mppi.cpp:331-334). NaN is absorbing in the sum,
        // so the total is the same without leaving the loop — and without a data-dependent branch at the head of every
        // step, behind which the scheduler cannot move the (independent) dynamics of the same step.
        total += sc;
        if (step + 1 == in.T) break;  // the state after the last step is never costed (mppi.cpp:316-341)
        // PinocchioDynamics::step, pinocchio_dynamics.cpp:226-260
        R tau[NJ], qdd[NJ], nle[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) { tau[i] = (i >= 3 && i < 10) ? u[i] : R(0); nle[i] = R(0); }
        if constexpr (FAITHFUL) {
            R sy, cy;
            sincos_(q[2], &sy, &cy);
            qd[0] = cy * u[0] - sy * u[1];
            qd[1] = sy * u[0] + cy * u[1];
            qd[2] = u[2];
            robot_calculate<R, true, POWER, KF>(M, q, qd, tau, qdd, nle, K);
        } else {
            // FUSED: qdd = M(q)^-1 tau through the structure-exploiting solver; joint sines / cosines are shared
            // (the loop-body build of the lean kernel evaluates them here, at the head of the solver: its objective has no yaw
            // term, and this placement gives the better schedule of the inertia loop; the unrolled build evaluates them after
            // the integration like the other objectives — no first-step test at the head of the loop, 59 modelled cycles and
            // the last spill less)
            if constexpr (LEAN && !BIG) { if (step != 0) joint_sincos<R>(F, q, cs, sn); }
            qd[0] = cs[2] * u[0] - sn[2] * u[1];
            qd[1] = sn[2] * u[0] + cs[2] * u[1];
            qd[2] = u[2];
            // (carrying the end effector point inside the INERTIA loop was tried: +4 registers, spills, 4 % slower)
            // (LEAN: the end effector point rides the solver's forward loop as a second, independent dependency chain)
            if constexpr (!LEAN) {   // the joint sines / cosines are shared
#if MPPI_ROLLED
                robot_calculate_rolled<R, POWER, KF>(M, q, qd, nle, K, cs, sn);
#else
                robot_calculate<R, false, POWER, KF, false, true>(M, q, qd, tau, qdd, nle, K, cs, sn);
#endif
            }
            // the objectives with kinematics (assisted manipulation, full reach-to-pose) always run the unrolled solver: with
            // the placements' structural zeros it executes 940 instructions per step fewer than the loop body (FP32 assisted
            // manipulation 5646 -> 4705, static model 6773 -> 5476 cycles), keeps its per-joint results in registers instead of
            // local memory, and adds 7 % to a step loop that is far beyond the instruction cache either way
            if constexpr (MIXED_SOLVER) {
                double tau64[NJ], qdd64[NJ];
#pragma unroll
                for (int i = 0; i < NJ; i++) tau64[i] = (i >= 3 && i < 10) ? u64[i] : 0.0;
                aba_fused_fast<double, MPPI_MIXED_SOLVER_UNROLL, false>(*F64, q64, cs64, sn64, tau64, qdd64);
                // base velocities from the control (pinocchio_dynamics.cpp:234-235), then semi-implicit Euler, all FP64
                qd64[0] = cs64[2] * u64[0] - sn64[2] * u64[1];
                qd64[1] = sn64[2] * u64[0] + cs64[2] * u64[1];
                qd64[2] = u64[2];
#pragma unroll
                for (int i = 0; i < NJ; i++) qd64[i] += qdd64[i] * in.dt64;
#pragma unroll
                for (int i = 0; i < NJ; i++) q64[i] += qd64[i] * in.dt64;
#pragma unroll
                for (int i = 0; i < NJ; i++) { q[i] = (R)q64[i]; qd[i] = (R)qd64[i]; }
            } else
            aba_fused_fast<R, (BIG || !LEAN) ? 7 : kArmUnroll, LEAN>(F, q, cs, sn, tau, qdd, &K.ee_pos);
        }
        if constexpr (MIXED_SOLVER) {
            if (POWER) {
                double p = 0.0;
#pragma unroll
                for (int i = 0; i < NJ; i++) p += (double)(tau[i] + nle[i]) * qd64[i];
                energy64 = std_max(0.0, energy64 + p * in.dt64);  // energy.hpp:19-22
                energy = (R)energy64;
            }
        } else if constexpr (MIXED) {
            // the base velocities were overwritten by the control (pinocchio_dynamics.cpp:234-235)
#pragma unroll
            for (int i = 0; i < 3; i++) qd64[i] = (double)qd[i];
#pragma unroll
            for (int i = 0; i < NJ; i++) qd64[i] += (double)qdd[i] * in.dt64;
#pragma unroll
            for (int i = 0; i < NJ; i++) q64[i] += qd64[i] * in.dt64;
#pragma unroll
            for (int i = 0; i < NJ; i++) { q[i] = (R)q64[i]; qd[i] = (R)qd64[i]; }
            if (POWER) {
                double p = 0.0;
#pragma unroll
                for (int i = 0; i < NJ; i++) p += (double)(tau[i] + nle[i]) * qd64[i];
                energy64 = std_max(0.0, energy64 + p * in.dt64);  // energy.hpp:19-22
                energy = (R)energy64;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NJ; i++) qd[i] += qdd[i] * in.dt;
#pragma unroll
            for (int i = 0; i < NJ; i++) q[i] += qd[i] * in.dt;
            if (POWER) {
                R p = R(0);
#pragma unroll
                for (int i = 0; i < NJ; i++) p += (tau[i] + nle[i]) * qd[i];
                energy = std_max(R(0), energy + p * in.dt);  // energy.hpp:19-22
            }
        }
        if constexpr (MIXED_SOLVER) {
            joint_sincos<double>(*F64, q64, cs64, sn64);
#pragma unroll
            for (int i = 2; i < 10; i++) { cs[i] = (R)cs64[i]; sn[i] = (R)sn64[i]; }
        } else if constexpr (!FAITHFUL && (!LEAN || BIG)) joint_sincos<R>(F, q, cs, sn);
#if defined(__CUDA_ARCH__)
        // the next iteration loads the noise of step + 2: its chunk is awaited HERE, at the loop's latch, where the step's
        // straight-line code ends anyway (a wait in front of the load split the step's basic block: +4 % kernel time)
        if constexpr (CHASE) { if ((step + 2) % CHASE_STEPS == 0 && step + 2 < in.T) noise_chase_wait<false>((step + 2) / CHASE_STEPS); }
#endif
    }
    return total;
}

// toy double integrator (BASELINE.json config 1)
template <class R>
MPPI_HD double rollout_toy(const ToyP<R> &P, const R *x0, const R *U, const R *eps, int T, R dt, double discount) {
    R px = x0[0], py = x0[1], vx = x0[2], vy = x0[3];
    double total = 0.0;
    for (int step = 0; step < T; ++step) {
        const R ux = U[step * 2] + eps[step * 2], uy = U[step * 2 + 1] + eps[step * 2 + 1];
        const R ex = px - P.target[0], ey = py - P.target[1];
        const R c = P.qp * (ex * ex + ey * ey) + P.qv * (vx * vx + vy * vy) + P.qu * (ux * ux + uy * uy);
        const double sc = discount_pow(discount, step) * (double)c;
        if (sc != sc) return sc;
        total += sc;
        vx += ux * dt; vy += uy * dt;
        px += vx * dt; py += vy * dt;
    }
    return total;
}

}  // namespace mppi_b200
