// K2 — the rollout kernel (replaces the thread-pool loop of reference src/controller/mppi.cpp:272-342
// and everything it calls per sample). One thread owns one rollout for the whole horizon: joint
// state, tank energy and stale kinematics live in registers, the shared control sequence / wrench
// table / initial state are staged once per block in shared memory, and the only per-step global
// traffic is the thread's own noise column (16-byte vector loads; the row is contiguous).
// After the horizon the block folds its costs into the running min / max (warp shuffles, then one
// order-preserving atomicMin / atomicMax per block) so no separate reduction pass is needed.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <math_constants.h>
#include "kernels.cuh"
#include "sample_core.cuh"

namespace mppi_b200 {

// The including translation unit defines MPPI_DEVICE_MODEL / MPPI_DEVICE_FAST_MODEL: its own
// __constant__ RobotModel<R> / FastModel<R> (one per precision, so no relocatable device code is needed).

// order preserving map double -> u64 (so that integer atomicMin/Max order like the doubles)
__device__ __forceinline__ unsigned long long encode_ordered(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double decode_ordered(unsigned long long e) {
    unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}

// Sharded rollout set (never batched): the LAST block of the rollout grid to get here publishes this rank's
// {-min, max, 0, valid slots} exchange payload and, with the peer-memory exchange attached, stores it into the peers'
// mailboxes — the first exchange of the update rides the rollout grid's tail instead of two launches of its own.
__device__ __forceinline__ void rollout_grid_epilogue(const DeviceState &d, const Frame *frame, int rollout_blocks) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(d.rollout_done, 1) == rollout_blocks - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (threadIdx.x == 0) {
            const int nvalid = atomicAdd(d.valid_count, 0);
            d.minmax_local[0] = nvalid > 0 ? -decode_ordered(atomicMin(&d.minmax_enc[0], ~0ull)) : -CUDART_INF;
            d.minmax_local[1] = nvalid > 0 ? decode_ordered(atomicMax(&d.minmax_enc[1], 0ull)) : -CUDART_INF;
            d.minmax_local[2] = nvalid >= 2 ? 2.0 : (double)nvalid;   // this rank's own count again: what the peer-memory exchange gathers (one element per rank)
            for (int r = 0; r < d.world; r++) d.minmax_local[3 + r] = (r == d.rank) ? (nvalid >= 2 ? 2.0 : (double)nvalid) : 0.0;
            *d.rollout_done = 0;
            __threadfence();
        }
        __syncthreads();
        if (d.has_px) exchange_push(d.px, EX_MINMAX, frame->attempt, d.minmax_local);
    }
}

// Two builds of the lean reach-to-pose variant, chosen by the launcher (measured on B200, FP64, update time):
//   BIG = false  one loop body for the seven arm joints, 168-register cap       K = 4096: 368 us   16384: 562   131072: 2260
//   BIG = true   arm joints unrolled, per-joint results in registers (255)      K = 4096: 398 us   16384: 461   131072: 2075
// (FP32: 297 / 330 / 1572 against 386 / 417 / 1577; the crossover sits near K = 12 k in FP64 and 24 k in FP32.)
// Few rollouts = one warp per SM, which lives on a small instruction footprint; many = fewer instructions win even at
// two blocks per SM. Those crossovers were measured when the unrolled step loop was 40 KB of code (3120 instructions per
// step) against 2600 for the loop body. Since the joint placements' structural zeros are dropped where the joint index is
// a constant (robot.cuh offset_mask / placement_is_flat), the unrolled step is 1927 instructions / 1501 FP64 against
// 2835 / 2095, its static model 4048 cycles against 4977 (FP32: 1756 against 3149), and its step loop is 31.0 KB — inside
// the 32 KB instruction cache level whose overflow cost the old unrolled build ~25 % at one warp per SM. The unrolled
// build therefore serves every rollout count by default; MPPI_B200_BIG_FROM=<rollouts> restores a crossover (12288 /
// 24576 were the measured ones) for A/B runs.
#ifndef MPPI_LEAN_MIN_BLOCKS
#define MPPI_LEAN_MIN_BLOCKS 1
#endif
template <class R, int VAR, bool FAITHFUL, class ParamsT, bool BIG = false, bool CHASE = false>
__global__ void __launch_bounds__(CHASE ? 256 : 128, (VAR == VAR_TP_LEAN && !FAITHFUL && !BIG) ? MPPI_LEAN_MIN_BLOCKS : 1) k_rollout(const __grid_constant__ DeviceState d, const __grid_constant__ ParamsT P, int optimal_only) {
    pdl_wait();
    // CHASE = noise chase (rollout_core.cuh), its own instantiation of the kernel: warp 0 of the block integrates 32 rollouts,
    // the block's other seven warps draw their noise
    constexpr bool chase = CHASE;
    // blockIdx.y = controller of a batched engine. Only the handful of buffers this kernel touches are offset
    // by hand (a full controller_view copy of the state costs ~40 registers here, i.e. a resident warp per SM).
    const size_t c = blockIdx.y;
    const size_t n = (size_t)d.nu * d.T;
    const Frame *frame = reinterpret_cast<const Frame *>(reinterpret_cast<const double *>(d.frame) + c * d.frame_doubles);
    const double *wrench = d.wrench + c * d.frame_doubles;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *sU = reinterpret_cast<R *>(smem_raw);        // nu*T
    R *sW = sU + d.nu * d.T;                         // 6*T
    R *sx = sW + 6 * d.T;                            // 32
    double *sDisc = reinterpret_cast<double *>(sx + 32);   // T discount factors std::pow(gamma, step) (mppi.cpp:326)
    const double *Usrc = d.U_shift + c * n;
    if (chase && frame->shift_by > 0) {
        // the prepare block (on the sampling warps of this grid's first block) is shifting the control sequence meanwhile: the same
        // shift of the published sequence, read directly (mppi.cpp:194-206)
        const long long shift = frame->shift_by, kept_steps = d.T - shift;
        for (int i = threadIdx.x; i < d.nu * d.T; i += blockDim.x) {
            const int t = i / d.nu, dd = i - t * d.nu;
            sU[i] = (R)d.U[(size_t)((t < kept_steps) ? (int)(t + shift) : d.T - 1) * d.nu + dd];
        }
    } else {
        for (int i = threadIdx.x; i < d.nu * d.T; i += blockDim.x) sU[i] = (R)Usrc[i];
    }
    const int has_w = frame->has_wrench;
    for (int i = threadIdx.x; i < 6 * d.T; i += blockDim.x) sW[i] = has_w ? (R)wrench[i] : R(0);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) sx[i] = (R)frame->x0[i];
    double *sx64 = sDisc + d.T;                            // 32: the initial state unrounded (FP32 fast mode)
    for (int i = threadIdx.x; i < 32; i += blockDim.x) sx64[i] = frame->x0[i];
    // the control sequence unrounded, for the FP64 state path of the fast mode (only those kernels reserve the room)
    constexpr bool STATE_PATH64 = sizeof(R) == 4 && MPPI_MIXED_STATE >= 2 && !FAITHFUL && (VAR == VAR_TP_FULL || VAR == VAR_AM || VAR == VAR_AM_ENERGY);
    double *sU64 = sx64 + 32;
    if constexpr (STATE_PATH64) for (int i = threadIdx.x; i < d.nu * d.T; i += blockDim.x) sU64[i] = Usrc[i];
    for (int i = threadIdx.x; i < d.T; i += blockDim.x) sDisc[i] = discount_pow(d.discount, i);
    if constexpr (CHASE) for (int i = threadIdx.x; i < (d.T + CHASE_STEPS - 1) / CHASE_STEPS; i += blockDim.x) s_chase_columns[i] = 0u;
    __syncthreads();

    // (noise chase: 32 rollouts per block whatever its size — the threads past warp 0 are the sampling threads)
    const long long k = chase ? (long long)blockIdx.x * 32 + threadIdx.x : (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double cost = 0.0;
    const bool active = optimal_only ? (k == 0) : (k < d.k_count && !(chase && threadIdx.x >= 32));
    if constexpr (CHASE) {
        if (chase && threadIdx.x >= 32) {
            const int j = (int)threadIdx.x - 32, ns = (int)blockDim.x - 32;
            if (blockIdx.x == 0) {
                // the update's prepare block first (it is short), flagged for the min / max atomics at the end of every block
                // (on the last sampling warp alone, so that this block's first chunk is not held up, it measured 1.6 us SLOWER)
                prepare_block(d, j, ns);
                asm volatile("bar.sync 1, %0;" ::"r"(ns) : "memory");
                if (j == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(d.chase_prepared), "r"((unsigned)(frame->attempt + 1ull)) : "memory"); }
            }
            chase_sampler<R, NJ>(d, d.Ldiag, (long long)blockIdx.x * 32, j, ns);
        }
    }
    if (active) {
        // the optimal re-rollout (Trajectory::filter, mppi.cpp:450-479) reads a row of zeros shared by all controllers
        const R *eps = static_cast<const R *>(d.noise) + (optimal_only ? (size_t)0 : (c * (size_t)d.k_count + (size_t)k) * n);
        if constexpr (VAR == VAR_TOY) {
            cost = rollout_toy<R>(P, sx, sU, eps, d.T, (R)d.dt, d.discount);
        } else {
            RolloutInputs<R> in;
            in.x0 = sx; in.x0_64 = sx64; in.U = sU; in.W = has_w ? sW : nullptr; in.T = d.T; in.dt = (R)d.dt; in.dt64 = d.dt; in.discount = d.discount; in.discount_table = sDisc;
            if constexpr (STATE_PATH64) in.U64 = sU64;
            double bd[7] = {0, 0, 0, 0, 0, 0, 0};
            cost = rollout_franka<R, VAR, FAITHFUL, ParamsT, BIG, CHASE>(MPPI_DEVICE_MODEL, MPPI_DEVICE_FAST_MODEL, P, in, eps, optimal_only ? bd : nullptr, MPPI_DEVICE_FAST_MODEL64);
            if (optimal_only && active) {
                for (int i = 0; i < 7; i++) d.breakdown[8 * c + i] = bd[i];
                d.breakdown[8 * c + 7] = cost;
            }
        }
        if (active) { if (optimal_only) d.optimal_cost[c] = cost; else d.costs[c * (size_t)d.k_count + k] = cost; }
    }
    if (optimal_only) return;
    if constexpr (CHASE) {
        if (chase) {   // the prepare block has reset the running min / max (long ago)
            const unsigned target = (unsigned)(frame->attempt + 1ull);
            unsigned seen;
            for (unsigned spins = 0;; spins++) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(d.chase_prepared) : "memory");
                if (seen == target) break;
                if (spins > (1u << 22)) __trap();
            }
        }
    }

    // block min / max over the non-NaN costs
    const bool valid = active && !(cost != cost);
    double mn = valid ? cost : __longlong_as_double(0x7ff0000000000000ll);   // +inf
    double mx = valid ? cost : __longlong_as_double(0xfff0000000000000ll);   // -inf
    int cnt = valid ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        atomicMin(&d.minmax_enc[2 * c], encode_ordered(mn));
        atomicMax(&d.minmax_enc[2 * c + 1], encode_ordered(mx));
        atomicAdd(d.valid_count + c, cnt);
    }
    if (d.world > 1) rollout_grid_epilogue(d, frame, (int)gridDim.x);
}

// ---- K2 for the FP32 fast mode of the objectives with kinematics (assisted manipulation, full reach-to-pose) ----------
// TWO WARPS PER 32 ROLLOUTS. With the state path in FP64 (rollout_core.cuh MIXED_SOLVER) a step is two computations
// that meet only in the state:
//   state warp  control + noise -> base velocity, joint sines / cosines, qdd = M(q)^-1 tau (the unrolled FP64 solver),
//               semi-implicit Euler                                        — needs nothing from the other warp
//   cost warp   single-precision kinematics + RNEA of that state, tank power and energy, the objective's stage cost,
//               the rollout's total                                         — consumes the states, one step behind
// As one warp they ran back to back through 84 KB of straight-line code per step, bound by instruction fetch (ncu, config
// 3: 51 % of the issue slots empty for lack of an instruction); as a producer and a consumer on two sub-partitions of
// the SM each runs its half, in parallel. States travel through a ring of shared-memory slots (lane-contiguous), handed
// over with named barriers (bar.arrive / bar.sync: the producer waits for the consumer only when the ring is full).
// Same arithmetic, same order as rollout_franka's MIXED_SOLVER path: costs are bit-identical (MPPI_B200_SPLIT=0 runs that
// path for A/B tests).
constexpr int SPLIT_DEPTH = 4;
struct SplitSlot {          // one step's state of 32 rollouts, lane fastest: every access is conflict-free
    double q[NJ][32];
    double qd[NJ][32];
    float cs[8][32], sn[8][32];   // joints 2..9
    float u[10][32];              // the control applied at this step (channels 10, 11 — the gripper — are ignored by the dynamics)
};
// (no fence in front of bar.arrive: the barrier orders the thread's earlier shared-memory accesses for its participants — PTX's
// own producer / consumer pattern — while a block-scope fence also waits for the state warp's noise loads in flight)
#ifndef MPPI_SPLIT_FENCE
#define MPPI_SPLIT_FENCE 0
#endif
__device__ __forceinline__ void split_arrive(int id) {
#if MPPI_SPLIT_FENCE
    __threadfence_block();
#endif
    asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void split_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

#if defined(MPPI_ROLLOUT_F32)
template <int VAR, class ParamsT>
__global__ void __launch_bounds__(64, 1) k_rollout_split(const __grid_constant__ DeviceState d, const __grid_constant__ ParamsT P) {
    typedef float R;
    constexpr int KF = VariantTraits<VAR>::kin;
    constexpr bool POWER = VariantTraits<VAR>::power;
    pdl_wait();
    const size_t c = blockIdx.y;
    const size_t n = (size_t)d.nu * d.T;
    const Frame *frame = reinterpret_cast<const Frame *>(reinterpret_cast<const double *>(d.frame) + c * d.frame_doubles);
    const double *wrench = d.wrench + c * d.frame_doubles;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SplitSlot *ring = reinterpret_cast<SplitSlot *>(smem_raw);
    double *sU64 = reinterpret_cast<double *>(ring + SPLIT_DEPTH);   // nu*T
    double *sDisc = sU64 + d.nu * d.T;                                // T
    double *sx64 = sDisc + d.T;                                       // 32
    R *sW = reinterpret_cast<R *>(sx64 + 32);                         // 6*T
    const double *Usrc = d.U_shift + c * n;
    for (int i = threadIdx.x; i < d.nu * d.T; i += blockDim.x) sU64[i] = Usrc[i];
    const int has_w = frame->has_wrench;
    for (int i = threadIdx.x; i < 6 * d.T; i += blockDim.x) sW[i] = has_w ? (R)wrench[i] : R(0);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) sx64[i] = frame->x0[i];
    for (int i = threadIdx.x; i < d.T; i += blockDim.x) sDisc[i] = discount_pow(d.discount, i);
    __syncthreads();

    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    long long k = (long long)blockIdx.x * 32 + lane;
    const bool active = k < d.k_count;
    if (!active) k = d.k_count - 1;   // both warps of a block walk the ring together: lanes past the end repeat the last rollout and write nothing
    const int T = d.T;
    const double dt64 = d.dt;
    if (role == 0) {
        // ---- state warp -------------------------------------------------------------------------------------------------
        const FastModel<double> &F64 = *MPPI_DEVICE_FAST_MODEL64;
        const R *eps = static_cast<const R *>(d.noise) + (c * (size_t)d.k_count + (size_t)k) * n;
        double q64[NJ], qd64[NJ], cs64[NJ], sn64[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) { q64[i] = sx64[i]; qd64[i] = sx64[NJ + i]; }
        joint_sincos<double>(F64, q64, cs64, sn64);
        R e_next[NJ];
        load_eps(eps, e_next);
        for (int step = 0; step < T; ++step) {
            double u64[NJ];
#pragma unroll
            for (int i = 0; i < NJ; i++) u64[i] = sU64[step * NJ + i] + (double)e_next[i];
            if (step + 1 < T) load_eps(eps + (step + 1) * NJ, e_next);
            const int s = step % SPLIT_DEPTH;
            if (step >= SPLIT_DEPTH) split_sync(1 + SPLIT_DEPTH + s);    // the cost warp has read what this slot held
            SplitSlot &slot = ring[s];
#pragma unroll
            for (int i = 0; i < NJ; i++) { slot.q[i][lane] = q64[i]; slot.qd[i][lane] = qd64[i]; }
#pragma unroll
            for (int i = 0; i < 8; i++) { slot.cs[i][lane] = (R)cs64[2 + i]; slot.sn[i][lane] = (R)sn64[2 + i]; }
#pragma unroll
            for (int i = 0; i < 10; i++) slot.u[i][lane] = (R)u64[i];
            split_arrive(1 + s);                                          // full
            if (step + 1 == T) break;
            double tau64[NJ], qdd64[NJ];
#pragma unroll
            for (int i = 0; i < NJ; i++) tau64[i] = (i >= 3 && i < 10) ? u64[i] : 0.0;
            aba_fused_fast<double, MPPI_MIXED_SOLVER_UNROLL, false>(F64, q64, cs64, sn64, tau64, qdd64);
            qd64[0] = cs64[2] * u64[0] - sn64[2] * u64[1];
            qd64[1] = sn64[2] * u64[0] + cs64[2] * u64[1];
            qd64[2] = u64[2];
#pragma unroll
            for (int i = 0; i < NJ; i++) qd64[i] += qdd64[i] * dt64;
#pragma unroll
            for (int i = 0; i < NJ; i++) q64[i] += qd64[i] * dt64;
            joint_sincos<double>(F64, q64, cs64, sn64);
        }
    } else {
        // ---- cost warp --------------------------------------------------------------------------------------------------
        const RobotModel<R> &M = MPPI_DEVICE_MODEL;
        double cost = 0.0;
        double energy64 = sx64[30];
        R energy = (R)energy64;
        Kinematics<R> K;
        R taunle[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) taunle[i] = R(0);
        // PinocchioDynamics::set_state -> calculate(): the kinematics the first stage cost reads, from the initial state (kept
        // out of the step loop: ~1600 instructions the loop would otherwise jump over every step)
        {
            R q0[NJ], qd0[NJ];
#pragma unroll
            for (int i = 0; i < NJ; i++) { q0[i] = (R)sx64[i]; qd0[i] = (R)sx64[NJ + i]; }
            robot_kinematics<R, KF>(M, q0, qd0, K);
        }
        for (int step = 0; step < T; ++step) {
            const int s = step % SPLIT_DEPTH;
            split_sync(1 + s);                                            // the state warp has filled this slot
            const SplitSlot &slot = ring[s];
            double q64[NJ], qd64[NJ];
            R q[NJ], qd[NJ], cs[NJ], sn[NJ], u[NJ];
#pragma unroll
            for (int i = 0; i < NJ; i++) { q64[i] = slot.q[i][lane]; qd64[i] = slot.qd[i][lane]; }
#pragma unroll
            for (int i = 0; i < 8; i++) { cs[2 + i] = slot.cs[i][lane]; sn[2 + i] = slot.sn[i][lane]; }
#pragma unroll
            for (int i = 0; i < 10; i++) u[i] = slot.u[i][lane];
            u[10] = u[11] = R(0); cs[0] = cs[1] = cs[10] = cs[11] = R(1); sn[0] = sn[1] = sn[10] = sn[11] = R(0);
            split_arrive(1 + SPLIT_DEPTH + s);                            // empty: the slot is in registers
#pragma unroll
            for (int i = 0; i < NJ; i++) { q[i] = (R)q64[i]; qd[i] = (R)qd64[i]; }
            if (POWER && step > 0) {   // tank: P = tau^T v with the velocities AFTER the step that produced this state (energy.hpp:19-22)
                double p = 0.0;
#pragma unroll
                for (int i = 0; i < NJ; i++) p += (double)taunle[i] * qd64[i];
                energy64 = std_max(0.0, energy64 + p * dt64);
                energy = (R)energy64;
            }
            const R yaw[2] = {cs[2], sn[2]};
            R cst;
            if constexpr (VAR == VAR_TP_FULL) cst = track_point_cost<R>(P, q, K, yaw, q64);
            else cst = assisted_cost<R>(P, q, qd, energy, K, has_w ? sW + step * 6 : nullptr, nullptr, yaw, q64, energy64);
            cost += sDisc[step] * (double)cst;
            if (step + 1 == T) break;
            R tau[NJ], qdd[NJ], nle[NJ];
#pragma unroll
            for (int i = 0; i < NJ; i++) { tau[i] = (i >= 3 && i < 10) ? u[i] : R(0); nle[i] = R(0); }
            qd[0] = cs[2] * u[0] - sn[2] * u[1];
            qd[1] = sn[2] * u[0] + cs[2] * u[1];
            qd[2] = u[2];
#if MPPI_ROLLED
            robot_calculate_rolled<R, POWER, KF>(M, q, qd, nle, K, cs, sn);
#else
            robot_calculate<R, false, POWER, KF, false, true>(M, q, qd, tau, qdd, nle, K, cs, sn);
#endif
#pragma unroll
            for (int i = 0; i < NJ; i++) taunle[i] = tau[i] + nle[i];
        }
        if (active) d.costs[c * (size_t)d.k_count + k] = cost;
        // block min / max over the non-NaN costs (this warp holds them)
        const bool valid = active && !(cost != cost);
        double mn = valid ? cost : __longlong_as_double(0x7ff0000000000000ll);
        double mx = valid ? cost : __longlong_as_double(0xfff0000000000000ll);
        int cnt = valid ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0 && cnt > 0) {
            atomicMin(&d.minmax_enc[2 * c], encode_ordered(mn));
            atomicMax(&d.minmax_enc[2 * c + 1], encode_ordered(mx));
            atomicAdd(d.valid_count + c, cnt);
        }
    }
    if (d.world > 1) rollout_grid_epilogue(d, frame, (int)gridDim.x);
}
#endif

template <class R, int VAR, bool FAITHFUL, class ParamsT>
cudaError_t launch_rollout_t(const DeviceState &d, const void *params, bool optimal_only, cudaStream_t s, int *chase_query) {
    const ParamsT &P = *static_cast<const ParamsT *>(params);
    // chase_query: no launch — whether the rollout blocks of a launch for this state would draw their own noise (engine creation)
    constexpr bool CAN_CHASE = VAR == VAR_TP_LEAN && !FAITHFUL;
    if (chase_query) { *chase_query = 0; if (!CAN_CHASE) return cudaSuccess; }
    // few rollouts: one warp per block spreads them over the SMs; many: 128-thread blocks
    int block = d.k_count <= 148 * 64 ? 32 : (d.k_count <= 148 * 256 ? 64 : 128);
    // the kernels with the 75 KB step body are bound by instruction fetch, which the warps of an SM share: 128-thread
    // blocks once there are enough rollouts to keep most SMs busy with them (config 3: 1225 -> 1200 us)
    if ((VAR == VAR_AM || VAR == VAR_AM_ENERGY || VAR == VAR_TP_FULL) && d.k_count >= 148 * 96) block = 128;
    long long grid = optimal_only ? 1 : (d.k_count + block - 1) / block;
    if (optimal_only) block = 32;
    size_t smem = sizeof(R) * ((size_t)d.nu * d.T + 6 * (size_t)d.T + 32) + sizeof(double) * ((size_t)d.T + 32);
    if (sizeof(R) == 4 && MPPI_MIXED_STATE >= 2 && !FAITHFUL && (VAR == VAR_TP_FULL || VAR == VAR_AM || VAR == VAR_AM_ENERGY)) smem += sizeof(double) * (size_t)d.nu * d.T;
    if constexpr (VAR == VAR_AM || VAR == VAR_AM_ENERGY || VAR == VAR_TP_FULL) {
        // A/B switch: block size of the one-warp kernels with the large step body. (Barriers inside their step loop were
        // measured too: the warps of an SM run this straight-line code in lockstep anyway — identical times to 0.1 us.)
        static const int env_block = std::getenv("MPPI_B200_AM_BLOCK") ? std::atoi(std::getenv("MPPI_B200_AM_BLOCK")) : 0;
        if (!optimal_only && env_block > 0) { block = env_block; grid = (d.k_count + block - 1) / block; }
    }
#if defined(MPPI_ROLLOUT_F32)
    if constexpr (sizeof(R) == 4 && MPPI_MIXED_STATE >= 2 && !FAITHFUL && (VAR == VAR_TP_FULL || VAR == VAR_AM || VAR == VAR_AM_ENERGY)) {
        static const bool split = !(std::getenv("MPPI_B200_SPLIT") && std::getenv("MPPI_B200_SPLIT")[0] == '0');
        // two warps per 32 rollouts pay when the doubled grid is still ONE wave (255 registers: four 64-thread blocks per SM);
        // beyond that — the 32 batched controllers of config 5 — the one-warp kernel's single wave is faster (measured 767 us
        // against 941)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const long long split_blocks = (d.k_count + 31) / 32 * (long long)d.batch;
        if (split && !optimal_only && split_blocks <= 4ll * sms) {
            const size_t ssmem = sizeof(SplitSlot) * SPLIT_DEPTH + sizeof(double) * ((size_t)d.nu * d.T + (size_t)d.T + 32) + sizeof(float) * 6 * (size_t)d.T;
            auto skern = k_rollout_split<VAR, ParamsT>;
            if (ssmem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
                if (e != cudaSuccess) return e;
            }
            return launch_level(rollout_overlap_level(split_blocks), skern, dim3((unsigned)((d.k_count + 31) / 32), d.batch), dim3(64), ssmem, s, d, P);
        }
    }
#endif
    auto kern = k_rollout<R, VAR, FAITHFUL, ParamsT, false>;
    if constexpr (VAR == VAR_TP_LEAN && !FAITHFUL) {
        static const long long big_from = std::getenv("MPPI_B200_BIG_FROM") ? std::atoll(std::getenv("MPPI_B200_BIG_FROM")) : 0;   // see the note above k_rollout
        if (d.k_count * (long long)d.batch >= big_from) kern = k_rollout<R, VAR, FAITHFUL, ParamsT, true>;   // the on-demand re-rollout runs the same build
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if constexpr (CAN_CHASE) {
        // Noise chase: a grid of one-warp blocks no larger than the machine — its blocks land one to an SM, which is otherwise
        // empty — gets seven sampling warps per block (256 threads x 255 registers: the SM's whole register file)
        if (chase_query) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const bool unrolled = kern == k_rollout<R, VAR, FAITHFUL, ParamsT, true>;   // (not the loop-body build of MPPI_B200_BIG_FROM)
            *chase_query = (unrolled && d.batch == 1 && block == 32 && grid <= sms && d.T <= CHASE_STEPS * CHASE_MAX_CHUNKS) ? 1 : 0;
            return cudaSuccess;
        }
        if (!optimal_only && d.chase) {
            auto ckern = k_rollout<R, VAR, FAITHFUL, ParamsT, true, true>;   // (the unrolled build)
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(ckern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
            }
            return launch_level(2, ckern, dim3((unsigned)grid, 1), dim3(256), smem, s, d, P, 0);
        }
    }
    return launch_level(rollout_overlap_level((long long)grid * d.batch), kern, dim3((unsigned)grid, d.batch), dim3(block), smem, s, d, P, optimal_only ? 1 : 0);
}

template <class R> cudaError_t launch_rollout_r(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query) {
    switch (variant) {
        case VAR_TOY: return launch_rollout_t<R, VAR_TOY, false, ToyP<R>>(d, params, optimal_only, s, chase_query);
        case VAR_TP_LEAN: return faithful ? launch_rollout_t<R, VAR_TP_LEAN, true, TrackPointP<R>>(d, params, optimal_only, s, chase_query)
                                          : launch_rollout_t<R, VAR_TP_LEAN, false, TrackPointP<R>>(d, params, optimal_only, s, chase_query);
        case VAR_TP_FULL: return faithful ? launch_rollout_t<R, VAR_TP_FULL, true, TrackPointP<R>>(d, params, optimal_only, s, chase_query)
                                          : launch_rollout_t<R, VAR_TP_FULL, false, TrackPointP<R>>(d, params, optimal_only, s, chase_query);
        case VAR_AM: return faithful ? launch_rollout_t<R, VAR_AM, true, AssistedP<R>>(d, params, optimal_only, s, chase_query)
                                     : launch_rollout_t<R, VAR_AM, false, AssistedP<R>>(d, params, optimal_only, s, chase_query);
        case VAR_AM_ENERGY: return faithful ? launch_rollout_t<R, VAR_AM_ENERGY, true, AssistedP<R>>(d, params, optimal_only, s, chase_query)
                                            : launch_rollout_t<R, VAR_AM_ENERGY, false, AssistedP<R>>(d, params, optimal_only, s, chase_query);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_rollout_f64(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query = nullptr);
cudaError_t launch_rollout_f32(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query = nullptr);

}  // namespace mppi_b200
