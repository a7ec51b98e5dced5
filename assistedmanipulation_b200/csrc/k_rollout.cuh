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
#include "kernels.cuh"

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
template <class R, int VAR, bool FAITHFUL, class ParamsT, bool BIG = false>
__global__ void __launch_bounds__(128, (VAR == VAR_TP_LEAN && !FAITHFUL && !BIG) ? MPPI_LEAN_MIN_BLOCKS : 1) k_rollout(const __grid_constant__ DeviceState d, const __grid_constant__ ParamsT P, int optimal_only, int lockstep) {
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
    for (int i = threadIdx.x; i < d.nu * d.T; i += blockDim.x) sU[i] = (R)Usrc[i];
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
    __syncthreads();

    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double cost = 0.0;
    const bool active = optimal_only ? (k == 0) : (k < d.k_count);
    // lockstep blocks meet at barriers inside the step loop: the threads past the end of the rollout set run the last
    // rollout once more (and write nothing) instead of leaving
    if (lockstep && !optimal_only && !active) k = d.k_count - 1;
    if (active || (lockstep && !optimal_only)) {
        // the optimal re-rollout (Trajectory::filter, mppi.cpp:450-479) reads a row of zeros shared by all controllers
        const R *eps = static_cast<const R *>(d.noise) + (optimal_only ? (size_t)0 : (c * (size_t)d.k_count + (size_t)k) * n);
        if constexpr (VAR == VAR_TOY) {
            cost = rollout_toy<R>(P, sx, sU, eps, d.T, (R)d.dt, d.discount);
        } else {
            RolloutInputs<R> in;
            in.x0 = sx; in.x0_64 = sx64; in.U = sU; in.W = has_w ? sW : nullptr; in.T = d.T; in.dt = (R)d.dt; in.dt64 = d.dt; in.discount = d.discount; in.discount_table = sDisc;
            in.lockstep = optimal_only ? 0 : lockstep;
            if constexpr (STATE_PATH64) in.U64 = sU64;
            double bd[7] = {0, 0, 0, 0, 0, 0, 0};
            cost = rollout_franka<R, VAR, FAITHFUL, ParamsT, BIG>(MPPI_DEVICE_MODEL, MPPI_DEVICE_FAST_MODEL, P, in, eps, optimal_only ? bd : nullptr, MPPI_DEVICE_FAST_MODEL64);
            if (optimal_only && active) {
                for (int i = 0; i < 7; i++) d.breakdown[8 * c + i] = bd[i];
                d.breakdown[8 * c + 7] = cost;
            }
        }
        if (active) { if (optimal_only) d.optimal_cost[c] = cost; else d.costs[c * (size_t)d.k_count + k] = cost; }
    }
    if (optimal_only) return;

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
}

template <class R, int VAR, bool FAITHFUL, class ParamsT>
cudaError_t launch_rollout_t(const DeviceState &d, const void *params, bool optimal_only, cudaStream_t s) {
    const ParamsT &P = *static_cast<const ParamsT *>(params);
    // few rollouts: one warp per block spreads them over the SMs; many: 128-thread blocks
    int block = d.k_count <= 148 * 64 ? 32 : (d.k_count <= 148 * 256 ? 64 : 128);
    // the kernels with the 75 KB step body are bound by instruction fetch, which the warps of an SM share: 128-thread
    // blocks once there are enough rollouts to keep most SMs busy with them (config 3: 1225 -> 1200 us)
    if ((VAR == VAR_AM || VAR == VAR_AM_ENERGY || VAR == VAR_TP_FULL) && d.k_count >= 148 * 96) block = 128;
    long long grid = optimal_only ? 1 : (d.k_count + block - 1) / block;
    if (optimal_only) block = 32;
    size_t smem = sizeof(R) * ((size_t)d.nu * d.T + 6 * (size_t)d.T + 32) + sizeof(double) * ((size_t)d.T + 32);
    if (sizeof(R) == 4 && MPPI_MIXED_STATE >= 2 && !FAITHFUL && (VAR == VAR_TP_FULL || VAR == VAR_AM || VAR == VAR_AM_ENERGY)) smem += sizeof(double) * (size_t)d.nu * d.T;
    int lockstep = 0;
    if constexpr (VAR == VAR_AM || VAR == VAR_AM_ENERGY || VAR == VAR_TP_FULL) {
        // experiment switches (A/B runs): block size and barriers per step of the kernels with the large step body
        static const int env_block = std::getenv("MPPI_B200_AM_BLOCK") ? std::atoi(std::getenv("MPPI_B200_AM_BLOCK")) : 0;
        static const int env_lockstep = std::getenv("MPPI_B200_LOCKSTEP") ? std::atoi(std::getenv("MPPI_B200_LOCKSTEP")) : 0;
        if (!optimal_only && env_block > 0) { block = env_block; grid = (d.k_count + block - 1) / block; }
        lockstep = env_lockstep;
    }
    auto kern = k_rollout<R, VAR, FAITHFUL, ParamsT, false>;
    if constexpr (VAR == VAR_TP_LEAN && !FAITHFUL) {
        static const long long big_from = std::getenv("MPPI_B200_BIG_FROM") ? std::atoll(std::getenv("MPPI_B200_BIG_FROM")) : 0;   // see the note above k_rollout
        if (d.k_count * (long long)d.batch >= big_from) kern = k_rollout<R, VAR, FAITHFUL, ParamsT, true>;   // the on-demand re-rollout runs the same build
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<dim3((unsigned)grid, d.batch), block, smem, s>>>(d, P, optimal_only ? 1 : 0, lockstep);
    return cudaGetLastError();
}

template <class R> cudaError_t launch_rollout_r(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s) {
    switch (variant) {
        case VAR_TOY: return launch_rollout_t<R, VAR_TOY, false, ToyP<R>>(d, params, optimal_only, s);
        case VAR_TP_LEAN: return faithful ? launch_rollout_t<R, VAR_TP_LEAN, true, TrackPointP<R>>(d, params, optimal_only, s)
                                          : launch_rollout_t<R, VAR_TP_LEAN, false, TrackPointP<R>>(d, params, optimal_only, s);
        case VAR_TP_FULL: return faithful ? launch_rollout_t<R, VAR_TP_FULL, true, TrackPointP<R>>(d, params, optimal_only, s)
                                          : launch_rollout_t<R, VAR_TP_FULL, false, TrackPointP<R>>(d, params, optimal_only, s);
        case VAR_AM: return faithful ? launch_rollout_t<R, VAR_AM, true, AssistedP<R>>(d, params, optimal_only, s)
                                     : launch_rollout_t<R, VAR_AM, false, AssistedP<R>>(d, params, optimal_only, s);
        case VAR_AM_ENERGY: return faithful ? launch_rollout_t<R, VAR_AM_ENERGY, true, AssistedP<R>>(d, params, optimal_only, s)
                                            : launch_rollout_t<R, VAR_AM_ENERGY, false, AssistedP<R>>(d, params, optimal_only, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_rollout_f64(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s);
cudaError_t launch_rollout_f32(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s);

}  // namespace mppi_b200
