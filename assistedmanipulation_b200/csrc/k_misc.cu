// K1 sample, K3 reduce/weights, K4 weighted sum, K5 finish — everything of
// mppi::Trajectory::sample / optimise (reference src/controller/mppi.cpp:189-270, 344-448) and the
// Savitzky–Golay window (src/controller/filter.cpp:19-173) that is not the rollout itself.
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"
#include "sample_core.cuh"

namespace mppi_b200 {

int rollout_overlap_level(long long blocks) {
    static const int sms = [] { int n = 148, dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    return blocks >= 2ll * sms ? 1 : 2;
}
int pdl_level() {
    static const int level = std::getenv("MPPI_B200_PDL") ? std::atoi(std::getenv("MPPI_B200_PDL")) : 1;
    return level;
}


cudaError_t upload_robot_model_f64();
cudaError_t upload_robot_model_f32();
cudaError_t upload_robot_model() {
    cudaError_t e = upload_robot_model_f64();
    if (e != cudaSuccess) return e;
    return upload_robot_model_f32();
}

__device__ __forceinline__ unsigned long long encode_ordered(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double decode_ordered(unsigned long long e) {
    unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}
cudaError_t launch_rollout_f64(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query = nullptr);
cudaError_t launch_rollout_f32(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query = nullptr);

// ---- warm start: the keep_best lowest-cost rollouts of the PREVIOUS update (stable order) -----------
// mppi.cpp:222-253. One block; each round finds the smallest (cost, index) pair that is strictly
// greater than the previous pick, which enumerates the stable sort order without sorting.
// Enumerates, in stable-sort order, the `keep` smallest (key, index) pairs of a candidate list: each round
// finds the smallest pair strictly greater than the previous pick. One block.
template <class GetKey, class GetIdx, class Emit>
__device__ __forceinline__ void select_smallest(long long count, long long keep, GetKey key_of, GetIdx idx_of, Emit emit) {
    __shared__ unsigned long long s_key[32];
    __shared__ long long s_idx[32];
    __shared__ unsigned long long last_key;
    __shared__ long long last_idx;
    if (threadIdx.x == 0) { last_key = 0; last_idx = -1; }
    __syncthreads();
    for (long long round = 0; round < keep; round++) {
        unsigned long long best = 0xffffffffffffffffull; long long bi = 0x7fffffffffffffffll;
        const unsigned long long lk = last_key; const long long li = last_idx;
        for (long long k = threadIdx.x; k < count; k += blockDim.x) {
            const long long gi = idx_of(k);
            if (gi < 2) continue;  // the two static rollouts are never candidates (mppi.cpp:222)
            const unsigned long long key = key_of(k);
            const bool after = (key > lk) || (key == lk && gi > li);
            if (after && (key < best || (key == best && gi < bi))) { best = key; bi = gi; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ok < best || (ok == best && oi < bi)) { best = ok; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { s_key[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int nw = blockDim.x >> 5;
            best = threadIdx.x < nw ? s_key[threadIdx.x] : 0xffffffffffffffffull;
            bi = threadIdx.x < nw ? s_idx[threadIdx.x] : 0x7fffffffffffffffll;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ok < best || (ok == best && oi < bi)) { best = ok; bi = oi; }
            }
            if (threadIdx.x == 0) { last_key = best; last_idx = bi; emit(round, best, bi); }
        }
        __syncthreads();
    }
}

// mppi.cpp:222-253 on this rank's rollouts. With one rank the picks ARE the kept set; with several the picks
// are this rank's candidates, all-gathered and merged by k_merge_kept.
__global__ void k_select_kept(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    for (long long k = threadIdx.x; k < d.k_count; k += blockDim.x) d.kept[k] = 0;
    __syncthreads();
    const long long keep = d.keep_best < d.K_total - 2 ? d.keep_best : d.K_total - 2;
    select_smallest(
        d.k_count, keep,
        [&](long long k) { double c = d.costs[k]; if (c != c) c = CUDART_INF; return encode_ordered(c); },  // NaN sorts last (SURVEY A-2)
        [&](long long k) { return k + d.k_begin; },
        [&](long long round, unsigned long long key, long long gi) {
            const bool found = gi != 0x7fffffffffffffffll;
            if (d.world == 1) {
                if (found) { d.kept[gi - d.k_begin] = 1; d.kept_list[round] = gi; }
            } else {
                d.cand[2 * round] = found ? decode_ordered(key) : CUDART_INF;
                d.cand[2 * round + 1] = __longlong_as_double(found ? gi : 0x7fffffffffffffffll);
            }
        });
}

// every rank merges the same world x keep candidate list into the global kept set
__global__ void k_merge_kept(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    const long long keep = d.keep_best < d.K_total - 2 ? d.keep_best : d.K_total - 2;
    select_smallest(
        (long long)d.world * keep, keep,
        [&](long long k) { return encode_ordered(d.cand_all[2 * k]); },
        [&](long long k) { const long long gi = __double_as_longlong(d.cand_all[2 * k + 1]); return gi == 0x7fffffffffffffffll ? -1 : gi; },
        [&](long long round, unsigned long long, long long gi) {
            if (gi == 0x7fffffffffffffffll) return;
            d.kept_list[round] = gi;
            if (gi >= d.k_begin && gi < d.k_begin + d.k_count) d.kept[gi - d.k_begin] = 1;
        });
}

// ---- K1 sample -----------------------------------------------------------------------------------
// One thread per (rollout, step) column. Fresh columns are eps = L z (z ~ N(0, I) from Philox, or the
// injected buffer); kept rollouts first move their surviving columns left by shift_by.
// In-place shift is safe because a kept row is processed by ONE block that stages it in shared memory.
template <class R, class RI, int NU> __device__ __forceinline__ void fresh_column(const DeviceState &d, const double *sL, long long kg, int t, R *dst) {
    constexpr int nu = NU;
    if (d.frame->noise_source != 0) {
        const RI *src = static_cast<const RI *>(d.injected) + ((size_t)(kg - d.k_begin) * d.T + t) * nu;
#pragma unroll
        for (int i = 0; i < nu; i++) dst[i] = (R)src[i];
        return;
    }
    float z[(NU + 3) / 4 * 4];
    const unsigned long long col = (unsigned long long)kg * (unsigned long long)d.T + (unsigned long long)t;
    const uint2 key = make_uint2((unsigned)d.frame->seed, (unsigned)(d.frame->seed >> 32));
    const unsigned upd = (unsigned)d.frame->update_index;
#pragma unroll
    for (int b = 0; b < (NU + 3) / 4; b++) {
        const uint4 r = philox4x32_10(make_uint4((unsigned)col, (unsigned)(col >> 32), (unsigned)b, upd), key);
        box_muller(r.x, r.y, &z[4 * b], &z[4 * b + 1]);
        box_muller(r.z, r.w, &z[4 * b + 2], &z[4 * b + 3]);
    }
    if (d.L_is_diagonal) {
#pragma unroll
        for (int i = 0; i < nu; i++) dst[i] = scale_noise(d.Ldiag[i], z[i], (R *)nullptr);   // constant indices: Ldiag stays in the parameter bank
        return;
    }
#pragma unroll
    for (int i = 0; i < nu; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < nu; j++) s += sL[j * nu + i] * (double)z[j];
        dst[i] = (R)s;
    }
}

template <class R, class RI, int NU> __global__ void __launch_bounds__(256) k_sample(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    // A block produces 256 consecutive columns = one contiguous span of 256*nu values: every thread
    // builds its column in shared memory, then the block streams the span out with 16-byte stores.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *tile = reinterpret_cast<R *>(smem_raw);
    __shared__ double sL[MAX_NU * MAX_NU];
    if (blockIdx.x == gridDim.x - 1) { prepare_block(d, (int)threadIdx.x, (int)blockDim.x); return; }
    if (!d.L_is_diagonal) { for (int i = threadIdx.x; i < d.nu * d.nu; i += blockDim.x) sL[i] = d.L[i]; __syncthreads(); }
    const long long col0 = (long long)blockIdx.x * blockDim.x;
    const long long col = col0 + threadIdx.x;   // local column index
    const long long ncols = d.k_count * d.T;
    constexpr int nu = NU;
    R *noise = static_cast<R *>(d.noise);
    if (col < ncols) {
        const long long kl = col / d.T;
        const int t = (int)(col - kl * d.T);
        const long long kg = kl + d.k_begin;
        R v[NU];
        if (kg == 0) {                 // rollout 0: zero noise, always (mppi.cpp:222)
#pragma unroll
            for (int i = 0; i < nu; i++) v[i] = R(0);
        } else if (kg == 1) {          // rollout 1 = -U_prev, the UNSHIFTED optimum (mppi.cpp:269)
#pragma unroll
            for (int i = 0; i < nu; i++) v[i] = (R)(-d.U[t * nu + i]);
        } else if (d.kept[kl]) {       // kept rollouts (k_shift_kept) keep their values
#pragma unroll
            for (int i = 0; i < nu; i++) v[i] = noise[(size_t)col * nu + i];
        } else {
            fresh_column<R, RI, NU>(d, sL, kg, t, v);
        }
#pragma unroll
        for (int i = 0; i < nu; i++) tile[threadIdx.x * nu + i] = v[i];
    }
    __syncthreads();
    const long long cols_here = (ncols - col0 < (long long)blockDim.x) ? (ncols - col0) : (long long)blockDim.x;
    const size_t span = (size_t)cols_here * nu;           // values in this block's span
    constexpr int VN = 16 / sizeof(R);
    R *dst = noise + (size_t)col0 * nu;
    if (((size_t)col0 * nu) % VN == 0) {
        const size_t nvec = span / VN;
        for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<int4 *>(dst)[i] = reinterpret_cast<const int4 *>(tile)[i];
        for (size_t i = nvec * VN + threadIdx.x; i < span; i += blockDim.x) dst[i] = tile[i];
    } else {
        for (size_t i = threadIdx.x; i < span; i += blockDim.x) dst[i] = tile[i];
    }
}

// K1 for a diagonal transform and NU % 4 == 0 (every Franka + Ridgeback configuration): one thread per Philox block
// = four consecutive values (sample_core.cuh), so a warp writes 32 consecutive quads straight from registers — no
// staging tile (its column-strided 8-byte stores were 4-way bank conflicts), no barrier, a third of the Philox /
// Box–Muller chain per thread, 32-bit index arithmetic, and kept rollouts are skipped instead of rewritten.
// Same counters and the same arithmetic as k_sample: the noise is bit-identical.
template <class R, class RI, int NU> __global__ void __launch_bounds__(256) k_sample_quads(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    if (blockIdx.x == gridDim.x - 1) { prepare_block(d, (int)threadIdx.x, (int)blockDim.x); return; }
    const long long quads = d.k_count * d.T * (NU / 4);
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= quads) return;
    long long kl; int t, b;
    quad_coordinates<NU>(g, quads, d.T, &kl, &t, &b);
    R v[4];
    if (!sample_quad<R, RI, NU>(d, dg.Ldiag, kl, t, b, v)) return;
    R *dst = static_cast<R *>(d.noise) + (size_t)g * 4;   // 16-byte aligned: the buffer is, and nu*T % 4 == 0
    if constexpr (sizeof(R) == 8) {
        reinterpret_cast<double2 *>(dst)[0] = make_double2((double)v[0], (double)v[1]);
        reinterpret_cast<double2 *>(dst)[1] = make_double2((double)v[2], (double)v[3]);
    } else {
        *reinterpret_cast<float4 *>(dst) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    }
}

// K1, one thread per COLUMN (NU values = NU/4 Philox blocks). The quad kernel above executed 246 instructions per thread
// for four values and was bound by exactly that (ncu, K = 131 072: 78 % issue active, 3.3 TB/s of writes in FP64 and
// the same 226 us for half the bytes in FP32); here the index arithmetic, the case analysis and the frame / kept reads
// are shared by the column's three blocks, the Philox products are 32 x 32 -> 64 multiplies, the square root is one
// MUFU and FP32 buffers are scaled in FP32. A thread stores its NU values as consecutive 16-byte vectors; a warp covers
// 32 consecutive columns = one contiguous span. Same counters as every other sampling path: bit-identical noise.
// (Eight resident blocks per SM — 32 registers, so that the 1026 blocks of config 2 are one wave instead of 1.16 — was measured:
// the same 14 us at config 2, 180 against 147 us at K = 131 072 in FP64, where the 40-register build spills nothing.)
template <class R, class RI, int NU> __global__ void __launch_bounds__(256) k_sample_columns(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    if (blockIdx.x == gridDim.x - 1) { prepare_block(d, (int)threadIdx.x, (int)blockDim.x); return; }
    const long long cols = d.k_count * d.T;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cols) return;
    long long kl; int t;
    column_coordinates(g, cols, d.T, &kl, &t);
    R v[NU];
    if (!sample_column<R, RI, NU>(d, dg.Ldiag, kl, t, v)) return;
    store_column<R, NU>(static_cast<R *>(d.noise) + (size_t)g * NU, v, (g & 1) != 0);
}

// kept rollouts: one block per kept rollout (mppi.cpp:243-252); nothing happens when shift_by <= 0
template <class R, class RI, int NU> __global__ void k_shift_kept(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *row = reinterpret_cast<R *>(smem_raw);
    __shared__ double sL[MAX_NU * MAX_NU];
    const long long shift = d.frame->shift_by;
    if (shift <= 0) return;
    for (int i = threadIdx.x; i < d.nu * d.nu; i += blockDim.x) sL[i] = d.L[i];
    const long long kg = d.kept_list[blockIdx.x];
    const long long kl = kg - d.k_begin;
    if (kl < 0 || kl >= d.k_count) return;
    const int n = d.nu * d.T;
    R *noise = static_cast<R *>(d.noise) + (size_t)kl * n;
    for (int e = threadIdx.x; e < n; e += blockDim.x) row[e] = noise[e];
    __syncthreads();
    const long long shifted = d.T - shift;
    for (int t = threadIdx.x; t < d.T; t += blockDim.x) {
        if (t < shifted) {
            for (int i = 0; i < d.nu; i++) noise[t * d.nu + i] = row[(t + shift) * d.nu + i];
        } else {
            R v[NU];
            fresh_column<R, RI, NU>(d, sL, kg, t, v);
#pragma unroll
            for (int i = 0; i < NU; i++) noise[t * NU + i] = v[i];
        }
    }
}

// ---- K3 ---------------------------------------------------------------------------------------------
// w_k = exp(-cost_scale (c_k - min) / (max - min)), NaN -> 0 (mppi.cpp:381-397); block partial sums
__global__ void __launch_bounds__(256) k_weights(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    __shared__ double s_part[8];
    double mm0, mm1, mm2;
    if (d.world == 1) {
        // single rank: no exchange, every block decodes the running min / max itself (no publish kernel)
        const int nvalid = *d.valid_count;
        mm0 = nvalid > 0 ? -decode_ordered(d.minmax_enc[0]) : -CUDART_INF;
        mm1 = nvalid > 0 ? decode_ordered(d.minmax_enc[1]) : -CUDART_INF;
        mm2 = nvalid >= 2 ? 2.0 : (double)nvalid;
        if (blockIdx.x == 0 && threadIdx.x == 0) { d.minmax[0] = mm0; d.minmax[1] = mm1; d.minmax[2] = mm2; }
    } else {
        // sharded: this rank's payload {-min, max, 0, valid slots} was published by the last block of the rollout grid.
        // Peer-memory exchange: wait for the peers' payloads in the local mailbox and combine in rank order (MAX; the
        // valid slots are disjoint, so MAX gathers them). NCCL / split ABI: the buffer was all-reduced (MAX) in place.
        const unsigned long long attempt = d.frame->attempt;
        // (the exchange description is read from the kernel's parameter bank, `dg`: indexing its arrays through the
        // controller view's copy would park the whole state in local memory; exchange_peer polls until the element is there)
        double valid = 0.0;
        if (dg.has_px) {
            // one thread per (rank, element) polls its word pair; a thread walking all peers one after the other paid an L2
            // round trip per peer (measured at 8 ranks: +6 us here, +25 us in k_finish)
            __shared__ double s_peer[3][MPPI_MAX_WORLD];
            if (threadIdx.x < 3 * d.world) {
                const int q = threadIdx.x / 3, e = threadIdx.x - 3 * q;
                s_peer[e][q] = (q == d.rank) ? d.minmax_local[e] : exchange_peer(dg.px, EX_MINMAX, attempt, q, e);   // element 2 = that rank's own valid count
            }
            __syncthreads();
            mm0 = -CUDART_INF; mm1 = -CUDART_INF;
            for (int q = 0; q < d.world; q++) { mm0 = fmax(mm0, s_peer[0][q]); mm1 = fmax(mm1, s_peer[1][q]); valid += s_peer[2][q]; }
        } else {
            mm0 = d.minmax_local[0]; mm1 = d.minmax_local[1];
            for (int q = 0; q < d.world; q++) valid += d.minmax_local[3 + q];
        }
        mm2 = valid >= 2.0 ? 2.0 : valid;   // saturate AFTER the sum: two ranks with one valid rollout each are two valid rollouts (mppi.cpp:368-370)
        if (blockIdx.x == 0 && threadIdx.x == 0) { d.minmax[0] = mm0; d.minmax[1] = mm1; d.minmax[2] = mm2; }
    }
    const double minimum = -mm0, maximum = mm1;
    const double difference = maximum - minimum;
    const bool bad = !(mm2 >= 2.0) || !(difference >= 1e-6);  // all-NaN / early return (mppi.cpp:368-375)
    if (bad) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *d.skip = 1;
    }
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double w = 0.0;
    if (k < d.k_count) {
        const double c = d.costs[k];
        if (c == minimum) atomicMin(d.argmin, k + d.k_begin);
        if (!bad) {
            w = (c != c) ? 0.0 : exp(-d.cost_scale * (c - minimum) / difference);
            d.weights[k] = w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += s_part[i];
        d.wsum_partial[blockIdx.x] = s;
    }
}

// ---- K4 weighted sum G = sum_k w_k eps_k (mppi.cpp:415-418) -----------------------------------------
// Memory bound: every noise element is read exactly once, 16 bytes per thread per load, rows are
// contiguous so a warp reads 512 contiguous bytes. Block b owns rollouts b, b+B, b+2B, ...; the
// partial sums are combined in a fixed order by k_gradient_reduce so results are reproducible.
template <class R> struct Vec16;
template <> struct Vec16<double> { typedef double2 type; static constexpr int n = 2; };
template <> struct Vec16<float> { typedef float4 type; static constexpr int n = 4; };
__device__ __forceinline__ void fma_vec(double *acc, double w, const double2 &v) { acc[0] = fma(w, v.x, acc[0]); acc[1] = fma(w, v.y, acc[1]); }
__device__ __forceinline__ void fma_vec(double *acc, double w, const float4 &v) {
    acc[0] = fma(w, (double)v.x, acc[0]); acc[1] = fma(w, (double)v.y, acc[1]); acc[2] = fma(w, (double)v.z, acc[2]); acc[3] = fma(w, (double)v.w, acc[3]);
}

// Row groups (G > 1): when a row has fewer 16-byte vectors than a block has threads (FP32 rows at T = 64: 192), the block's
// threads split into G groups that walk different rows, so every thread has loads in flight (the FP32 kernel ran 192 of
// 256 threads and reached 2.3 TB/s where the FP64 one reaches 5.7); the groups' sums meet in shared memory in group order.
// G is a template parameter: G = 1 is the plain kernel, whose row loop the compiler unrolls on its own (128 registers).
// FUSED (one rank, at most MPPI_FUSED_ROWS rollouts per block): the block first computes the weights of ITS rollouts — what
// k_weights does for all of them (mppi.cpp:381-397), including the decode of the running min / max, the skip decision and
// the argmin — keeps them in shared memory, and leaves its partial sum of the weights in wsum_partial[block]; its row of
// partial sums is written channel-major ([nu][T]) for k_finish, which combines the rows itself. Two kernels (k_weights,
// k_gradient_reduce) and their launch gaps less per update: ~6 us of a 200 us update at K = 4096.
template <class R, int G, bool FUSED> __global__ void __launch_bounds__(512) k_gradient(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    typedef typename Vec16<R>::type V;
    constexpr int VN = Vec16<R>::n;
    extern __shared__ __align__(16) double s_group[];   // (G - 1) x n partial sums
    __shared__ double s_w[FUSED ? MPPI_FUSED_ROWS : 1];
    const long long stride = (long long)gridDim.x * G;
    if constexpr (FUSED) {
        __shared__ double s_part[16];
        const int nvalid = *d.valid_count;
        const double mm0 = nvalid > 0 ? -decode_ordered(d.minmax_enc[0]) : -CUDART_INF;
        const double mm1 = nvalid > 0 ? decode_ordered(d.minmax_enc[1]) : -CUDART_INF;
        const double mm2 = nvalid >= 2 ? 2.0 : (double)nvalid;
        const double minimum = -mm0, difference = mm1 - minimum;
        const bool bad = !(mm2 >= 2.0) || !(difference >= 1e-6);  // all-NaN / early return (mppi.cpp:368-375)
        if (blockIdx.x == 0 && threadIdx.x == 0) { d.minmax[0] = mm0; d.minmax[1] = mm1; d.minmax[2] = mm2; if (bad) *d.skip = 1; }
        // local row li = j * G + g  <->  rollout (blockIdx.x * G + g) + j * stride  (group g walks j = 0, 1, ...)
        double w_sum = 0.0;
        for (int li = threadIdx.x; li < MPPI_FUSED_ROWS; li += blockDim.x) {
            const long long k = (long long)blockIdx.x * G + (li % G) + (long long)(li / G) * stride;
            double w = 0.0;
            if (k < d.k_count) {
                const double c = d.costs[k];
                if (c == minimum) atomicMin(d.argmin, k + d.k_begin);
                if (!bad) { w = (c != c) ? 0.0 : exp(-d.cost_scale * (c - minimum) / difference); d.weights[k] = w; }
            }
            s_w[li] = w;
            w_sum += w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w_sum += __shfl_xor_sync(0xffffffffu, w_sum, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = w_sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < (int)((blockDim.x + 31) >> 5); i++) t += s_part[i];   // fixed order
            d.wsum_partial[blockIdx.x] = t;
        }
        if (bad) return;
    } else {
        if (*d.skip) return;
    }
    const int n = d.nu * d.T;
    const int nvec = n / VN;  // host guarantees divisibility
    const V *noise = static_cast<const V *>(d.noise);
    const int g = G > 1 ? (int)threadIdx.x / nvec : 0;
    const int lanes = G > 1 ? nvec : (int)blockDim.x;          // threads that walk a row together
    const long long first = (long long)blockIdx.x * G + g;
    // weight of this thread's j-th row (rollout first + j * stride)
    auto weight_of = [&](long long k, int j) -> double { if constexpr (FUSED) return s_w[j * G + g]; else return d.weights[k]; };
    // where element i of the row goes in the block's row of partial sums
    auto slot_of = [&](int i) -> size_t { if constexpr (FUSED) { const int t = i / d.nu; return (size_t)(i - t * d.nu) * d.T + t; } else return (size_t)i; };
    for (int e = G > 1 ? (int)threadIdx.x - g * nvec : (int)threadIdx.x; e < nvec && g < G; e += lanes) {
        double acc[VN];
#pragma unroll
        for (int i = 0; i < VN; i++) acc[i] = 0.0;
        long long k = first;
        int j = 0;
        // FP32 rows: 8 rows in flight per thread (a 16-byte load carries half as many rows' worth of latency-hiding bytes per
        // element as in FP64, and the register budget allows it: one 384-thread block per SM either way)
        if (sizeof(R) == 4) {
            for (; k + 7 * stride < d.k_count; k += 8 * stride, j += 8) {
                V v[8]; double w[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = __ldg(noise + (size_t)(k + q * stride) * nvec + e);
#pragma unroll
                for (int q = 0; q < 8; q++) w[q] = weight_of(k + q * stride, j + q);
#pragma unroll
                for (int q = 0; q < 8; q++) fma_vec(acc, w[q], v[q]);
            }
        }
        // 4 rows in flight per thread
        for (; k + 3 * stride < d.k_count; k += 4 * stride, j += 4) {
            const V v0 = __ldg(noise + (size_t)k * nvec + e);
            const V v1 = __ldg(noise + (size_t)(k + stride) * nvec + e);
            const V v2 = __ldg(noise + (size_t)(k + 2 * stride) * nvec + e);
            const V v3 = __ldg(noise + (size_t)(k + 3 * stride) * nvec + e);
            const double w0 = weight_of(k, j), w1 = weight_of(k + stride, j + 1), w2 = weight_of(k + 2 * stride, j + 2), w3 = weight_of(k + 3 * stride, j + 3);
            fma_vec(acc, w0, v0); fma_vec(acc, w1, v1); fma_vec(acc, w2, v2); fma_vec(acc, w3, v3);
        }
        for (; k < d.k_count; k += stride, j++) fma_vec(acc, weight_of(k, j), __ldg(noise + (size_t)k * nvec + e));
        if (G == 1 || g == 0) {
#pragma unroll
            for (int i = 0; i < VN; i++) d.grad_partial[(size_t)blockIdx.x * n + (G == 1 ? slot_of(e * VN + i) : (size_t)(e * VN + i))] = acc[i];
        } else {
#pragma unroll
            for (int i = 0; i < VN; i++) s_group[(size_t)(g - 1) * n + e * VN + i] = acc[i];
        }
    }
    if (G > 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double sum = d.grad_partial[(size_t)blockIdx.x * n + i];
            for (int q = 1; q < G; q++) sum += s_group[(size_t)(q - 1) * n + i];
            if constexpr (!FUSED) d.grad_partial[(size_t)blockIdx.x * n + i] = sum;
            else s_group[(size_t)(G - 1) * n + i] = sum;   // (fused: the row is rewritten channel-major below, through a staging row)
        }
        if constexpr (FUSED) {
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) d.grad_partial[(size_t)blockIdx.x * n + slot_of(i)] = s_group[(size_t)(G - 1) * n + i];
        }
    }
}

// fixed-order combination of the partials into the exchange buffer sums = {sum w, sum w*eps}.
// Block = 32 elements x 32 slices of the partial rows (every thread has at most a few loads, all in
// flight at once); slice sums are combined in slice order.
__global__ void __launch_bounds__(1024) k_gradient_reduce(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    __shared__ double part[32][33];
    // A skipped update (max - min < 1e-6 or fewer than two valid rollouts, mppi.cpp:368-375) leaves weights and gradient
    // untouched, but a sharded set still exchanges this buffer (the argmin slots ride it, and every rank must find its
    // peers' flags): it is rewritten on EVERY update — zeros for the sums — so nothing stale is ever combined twice.
    const int skip = *d.skip;
    if (skip && d.world == 1) return;
    const int n = d.nu * d.T;
    const int rows = d.grad_blocks;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (e < n && !skip) {
        int b = slice;
        for (; b + 96 < rows; b += 128) {
            s0 += d.grad_partial[(size_t)b * n + e];
            s1 += d.grad_partial[(size_t)(b + 32) * n + e];
            s2 += d.grad_partial[(size_t)(b + 64) * n + e];
            s3 += d.grad_partial[(size_t)(b + 96) * n + e];
        }
        for (; b < rows; b += 32) s0 += d.grad_partial[(size_t)b * n + e];
    }
    part[slice][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (slice == 0 && e < n) {
        double s = part[0][lane];
#pragma unroll
        for (int i = 1; i < 32; i++) s += part[i][lane];
        d.sums[1 + e] = s;
    }
    if (blockIdx.x == 0 && slice == 1) {
        // sum of the weights: warp-strided partial sums, then a fixed-order shuffle tree
        double s = 0.0;
        if (!skip) for (int b = lane; b < d.weight_blocks; b += 32) s += d.wsum_partial[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            d.sums[0] = s;
            // sharded: one slot per rank after {sum w, sum w*eps} carries this rank's best global index + 1
            // (0 = the global minimum is not here), so the SUM exchange doubles as an all-gather of the argmin
            if (d.world > 1) {
                const long long a = *d.argmin;
                for (int r = 0; r < d.world; r++) d.sums[1 + d.nu * d.T + r] = (r == d.rank && a != 0x7fffffffffffffffll) ? (double)(a + 1) : 0.0;
            }
        }
    }
    if (d.world > 1 && dg.has_px) {
        // the LAST block to get here stores this rank's {sum w, sum w*eps, argmin slots} into the peers' mailboxes:
        // the second exchange of the update rides this kernel's tail (k_finish combines the slots)
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(d.reduce_done, 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (threadIdx.x == 0) *d.reduce_done = 0;
            exchange_push(dg.px, EX_SUMS, d.frame->attempt, d.sums);
        }
    }
}

// ---- K5 finish: gradient step, Savitzky–Golay smoothing, clamp, publish ------------------------------
// mppi.cpp:415-447 + filter.cpp:19-173. One block; channel d of the window is run by thread d with the
// window staged in shared memory (the recurrence is sequential in time, channels are independent).
__device__ __forceinline__ int sg_lower_bound(const double *tt, int len, double t) {  // std::lower_bound
    int first = 0;
    while (len > 0) {
        const int half = len >> 1, mid = first + half;
        if (tt[mid] < t) { first = mid + 1; len = len - half - 1; } else len = half;
    }
    return first;
}

// The recurrent part of the Savitzky-Golay application in registers (transposed form): when res_i is known it is
// added, with its tap, to the NR = w-1 accumulators of the applications that will read it; every step is NR
// independent FMAs and the chain from res_i to res_{i+1} is a single FMA.
template <int NR> __device__ __forceinline__ void sg_recurrence(const double *F, const double *sw, int T, double *u, int w, double *sU) {
    double c[NR], acc[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) { c[j] = sw[j]; acc[j] = j < T ? F[j] : 0.0; }
    for (int i = 0; i < T; i++) {
        const double res = acc[0];
        u[w + i - 1] = res; sU[i] = res;
        const double next = i + NR < T ? F[i + NR] : 0.0;
#pragma unroll
        for (int m = 1; m < NR; m++) acc[m - 1] = fma(c[NR - m], res, acc[m]);
        acc[NR - 1] = fma(c[0], res, next);
    }
}
__device__ __forceinline__ bool sg_recurrence_dispatch(int nr, const double *F, const double *sw, int T, double *u, int w, double *sU) {
    switch (nr) {
        case 1: sg_recurrence<1>(F, sw, T, u, w, sU); return true;
        case 2: sg_recurrence<2>(F, sw, T, u, w, sU); return true;
        case 3: sg_recurrence<3>(F, sw, T, u, w, sU); return true;
        case 4: sg_recurrence<4>(F, sw, T, u, w, sU); return true;
        case 5: sg_recurrence<5>(F, sw, T, u, w, sU); return true;
        case 7: sg_recurrence<7>(F, sw, T, u, w, sU); return true;
        case 9: sg_recurrence<9>(F, sw, T, u, w, sU); return true;    // the default window of 10
        case 14: sg_recurrence<14>(F, sw, T, u, w, sU); return true;
        case 19: sg_recurrence<19>(F, sw, T, u, w, sU); return true;
        default: return false;
    }
}

// staged: the block's shared-memory copy {sum w (combined), argmin slots} when the peer exchange is attached (k_finish)
__device__ __forceinline__ void finish_publish_stats(const DeviceState &d, const PeerExchange *px, int n, const double *s_total = nullptr, const double *s_arg = nullptr) {
    d.result[n + 0] = d.minmax[0]; d.result[n + 1] = d.minmax[1]; d.result[n + 2] = d.minmax[2];
    long long best = *d.argmin;
    if (d.world > 1) {   // lowest global index among the ranks that hold the global minimum (mppi.cpp:363-366 order)
        best = 0x7fffffffffffffffll;
        for (int r = 0; r < d.world; r++) { const double v = px ? s_arg[r] : d.sums[1 + n + r]; if (v > 0.0 && (long long)v - 1 < best) best = (long long)v - 1; }
    }
    d.result[n + 3] = __longlong_as_double(best);
    d.result[n + 4] = s_total ? *s_total : d.sums[0];
    // Last word of the host-mapped block: the number of this update. The host polls it instead of waiting for the
    // stream's end-of-update event (engine.cu: wait_published) — the block is complete when the word arrives (the fence
    // orders every earlier store of this block's threads, which met at a barrier, ahead of it; PCIe keeps posted writes in order).
    __threadfence_system();
    *reinterpret_cast<volatile double *>(d.result + n + 7) = (double)(d.frame->attempt + 1ull);
}

__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ DeviceState dg) {
    pdl_wait();
    const DeviceState d = controller_view(dg, blockIdx.y);
    // One BLOCK per control channel (channels are independent, mppi.cpp:424-447): the elementwise work runs over the
    // channel's T entries, the window recurrence on warp 0 with its own scheduler.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sU = reinterpret_cast<double *>(smem_raw);   // T working values of this channel
    double *u = sU + d.T;                                  // Lw window samples
    double *tm = u + d.sg_len;                             // Lw window times
    double *sw = tm + d.sg_len;                            // 2w+1 taps
    const int n = d.nu * d.T, ch = blockIdx.x;
    const int skip = *d.skip;
    const PeerExchange *px = (d.world > 1 && dg.has_px) ? &dg.px : nullptr;   // in the parameter bank (see k_weights)
    // Peer exchange: the block's threads poll the peers' words of THIS channel's T elements and of {sum w} in parallel — one
    // (rank, element) per thread and pass — and leave the rank-ordered sums in shared memory.
    double *s_comb = sw + 2 * d.sg_window + 1 + d.T;   // T + 1 combined values: [0] = sum w, [1 + t] = channel element t
    double *s_arg = s_comb + d.T + 1;                   // world argmin slots: slot r is filled by rank r alone (the others add 0)
    if (px) {
        const unsigned long long attempt = d.frame->attempt;
        const int per_rank = d.T + 1, items = per_rank * (d.world - 1);
        double *s_raw = s_arg + d.world;                // (world - 1) x (T + 1) peers' values
        for (int i = threadIdx.x; i < items + d.world; i += blockDim.x) {
            if (i >= items) {                           // rank r's own argmin slot (one poll per rank; the publishing thread walking all world x world slots cost 18 us at 8 ranks)
                const int r = i - items;
                s_arg[r] = (r == d.rank) ? d.sums[1 + n + r] : exchange_peer(*px, EX_SUMS, attempt, r, 1 + n + r);
                continue;
            }
            const int slot = i / per_rank, j = i - slot * per_rank;
            const int q = slot < d.rank ? slot : slot + 1;
            s_raw[i] = exchange_peer(*px, EX_SUMS, attempt, q, j == 0 ? 0 : 1 + (j - 1) * d.nu + ch);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < per_rank; j += blockDim.x) {
            const double own = d.sums[j == 0 ? 0 : 1 + (j - 1) * d.nu + ch];
            double acc = 0.0;
            for (int q = 0; q < d.world; q++) acc += (q == d.rank) ? own : s_raw[(q < d.rank ? q : q - 1) * per_rank + j];   // rank order: identical bits on every rank
            s_comb[j] = acc;
        }
        __syncthreads();                               // ... and a peer that never arrived is known to every thread below
    }
    // Fused tail (one rank): this block combines its channel's T elements over the rows of partial sums the weighted-sum
    // kernel left channel-major — coalesced 8 T-byte reads, `slices` rows in flight per element, fixed order — and the blocks'
    // partial sums of the weights; k_gradient_reduce is not launched.
    const bool fused = d.fused_tail != 0;
    if (fused) {
        double *s_red = s_comb + d.T + 1;
        const int slices = min(8, max(1, (int)blockDim.x / d.T)), rows = d.grad_blocks;
        for (int i = threadIdx.x; i < slices * d.T; i += blockDim.x) {
            const int sl = i / d.T, t = i - sl * d.T;
            const double *src = d.grad_partial + (size_t)ch * d.T + t;
            double a[8];
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] = 0.0;
            int b = sl;
            for (; b + 7 * slices < rows; b += 8 * slices) {   // eight rows in flight per thread
#pragma unroll
                for (int q = 0; q < 8; q++) a[q] += src[(size_t)(b + q * slices) * n];
            }
            for (; b < rows; b += slices) a[0] += src[(size_t)b * n];
            s_red[i] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        }
        if (threadIdx.x >= blockDim.x - 32) {   // the last warp: sum of the weights, warp-strided partial sums and a fixed-order shuffle tree
            const int lane = threadIdx.x & 31;
            double w = 0.0;
            for (int b = lane; b < rows; b += 32) w += d.wsum_partial[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if (lane == 0) s_comb[0] = w;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < d.T; t += blockDim.x) {
            double a = s_red[t];
            for (int sl = 1; sl < slices; sl++) a += s_red[sl * d.T + t];
            s_comb[1 + t] = a;
        }
        __syncthreads();
    }
    const bool staged = px || fused;
    const double total = staged ? s_comb[0] : d.sums[0];
    // "all nan rollouts" (mppi.cpp:368-370): the reference throws before publishing. A peer that never arrived (exchange
    // time-out) is handled the same way: the engine keeps its last good control sequence, the host reports the error.
    const bool dead = !(d.minmax[2] >= 2.0) || (px && *px->error_dev);
    if (dead) {   // nothing is published but the statistics
        if (ch == 0 && threadIdx.x == 0) finish_publish_stats(d, px, n, staged ? s_comb : nullptr, s_arg);
        return;
    }
    for (int t = threadIdx.x; t < d.T; t += blockDim.x) {
        const int e = t * d.nu + ch;
        double v = d.U_shift[e];
        if (!skip) {
            const double g = (staged ? s_comb[1 + t] : d.sums[1 + e]) / total;   // weights are normalised by the total (mppi.cpp:403-408)
            d.gradient[e] = g;
            v += g * d.gradient_step;
        }
        sU[t] = v;
    }
    if (!skip && d.sg_enabled) {
        const int Lw = d.sg_len, w = d.sg_window, ntaps = 2 * w + 1;
        for (int i = threadIdx.x; i < Lw; i += blockDim.x) { u[i] = d.sg_uu[ch * Lw + i]; tm[i] = d.sg_tt[ch * Lw + i]; }
        for (int i = threadIdx.x; i < ntaps; i += blockDim.x) sw[i] = d.sg_weights[i];
        __syncthreads();
        // Warp 0: the window bookkeeping is lane-parallel, only the T applications are sequential (each reads the
        // value the previous one wrote), and each of those is a lane-parallel dot product over the 2w+1 taps.
        const int lane = threadIdx.x & 31;
        if (threadIdx.x < 32) {
            const double t0 = d.frame->time;
            // trim(t0): filter.cpp:34-67. start_idx is w before the first pass and w + T after any pass.
            const int start_idx = d.sg_started[ch] ? w + d.T : w;
            int trim_idx = start_idx;
            for (int i = lane; i < start_idx; i += 32) if (tm[i] >= t0) { trim_idx = i; break; }  // first hit of this lane
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) trim_idx = min(trim_idx, __shfl_xor_sync(0xffffffffu, trim_idx, o));
            const int offset = trim_idx - w;
            if (offset > 0) {
                // rotate left by offset, then refill the vacated tail with the last valid sample
                const double lu = u[Lw - 1], lt = tm[Lw - 1];  // element that lands at Lw - offset - 1
                for (int base = 0; base < Lw; base += 32) {
                    const int i = base + lane;
                    const bool in = i + offset < Lw;
                    const double a = in ? u[i + offset] : lu, b = in ? tm[i + offset] : lt;
                    __syncwarp();
                    if (i < Lw) { u[i] = a; tm[i] = b; }
                    __syncwarp();
                }
            }
            if (lane == 0) tm[w] = t0;
            __syncwarp();
            // add_measurement x T (filter.cpp:69-90): final state = samples in [w, w+T), copies of the last beyond.
            // Times are compared with == / >= later: no FMA contraction, exactly m_rollout_time + i * m_time_step (mppi.cpp:430).
            for (int i = lane; i < Lw - w; i += 32) {
                const int k = i < d.T ? i : d.T - 1;
                u[w + i] = sU[k];
                tm[w + i] = __dadd_rn(t0, __dmul_rn((double)k, d.dt));
            }
            __syncwarp();
            // apply x T (filter.cpp:163-173): the filtered value is written ONE SLOT EARLIER than the sample, so
            // application i reads what applications i-w+1 .. i-1 wrote (slots i .. i+w-2 of its 2w+1 taps) and
            // untouched samples beyond: an order-(w-1) recurrence plus a feed-forward part.
            // Fast path (every lookup lands on its own sample, idx_i = w + i — always, unless the window holds
            // duplicate times): the feed-forward part F_i = sum_{j>=w-1} sw[j] u[i+j] is evaluated for all i by
            // the lanes at once from the untouched window; lane 0 then runs the short recurrence in place.
            bool regular = true;
            for (int i = lane; i < d.T; i += 32) {
                const double t = __dadd_rn(t0, __dmul_rn((double)i, d.dt));
                if (!(tm[w + i] >= t && tm[w + i - 1] < t)) regular = false;
            }
            regular = __all_sync(0xffffffffu, regular) && w >= 1;
            if (regular) {
                double *F = sw + ntaps;   // T feed-forward sums
                for (int i = lane; i < d.T; i += 32) {
                    double acc = 0.0;
                    for (int j = w - 1; j < ntaps; j++) acc = fma(sw[j], u[i + j], acc);
                    for (int j = 0; j < w - 1 - i; j++) acc = fma(sw[j], u[i + j], acc);   // history left of the first written slot
                    F[i] = acc;
                }
                __syncwarp();
                if (lane == 0 && !sg_recurrence_dispatch(w - 1, F, sw, d.T, u, w, sU)) {
                    for (int i = 0; i < d.T; i++) {   // any other window: the same recurrence through shared memory
                        double acc = F[i];
                        for (int j = (w - 1 - i > 0 ? w - 1 - i : 0); j < w - 1; j++) acc = fma(sw[j], u[i + j], acc);
                        u[w + i - 1] = acc; sU[i] = acc;
                    }
                }
                __syncwarp();
            } else {
            const double my_w0 = lane < ntaps ? sw[lane] : 0.0;
            for (int i = 0; i < d.T; i++) {
                const double t = __dadd_rn(t0, __dmul_rn((double)i, d.dt));
                // std::lower_bound over the (sorted) times; the common answer w + i is verified, else searched
                int idx = w + i;
                if (!(tm[idx] >= t && tm[idx - 1] < t)) idx = sg_lower_bound(tm, Lw, t);
                const double *v = u + idx - w;
                double res = lane < ntaps ? my_w0 * v[lane] : 0.0;
                for (int j = lane + 32; j < ntaps; j += 32) res = fma(sw[j], v[j], res);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
                __syncwarp();
                if (lane == 0) { u[idx - 1] = res; sU[i] = res; }
                __syncwarp();
            }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < Lw; i += blockDim.x) { d.sg_uu[ch * Lw + i] = u[i]; d.sg_tt[ch * Lw + i] = tm[i]; }
        if (threadIdx.x == 0) d.sg_started[ch] = 1;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < d.T; t += blockDim.x) {
        const int e = t * d.nu + ch;
        double v = sU[t];
        if (!skip && d.bound) {
            v = std_min(v, dg.cmax[ch]);   // (dg: dynamic index into the parameter bank, not into a local copy of the state)   // cwiseMin(max).cwiseMax(min), mppi.cpp:443-447; std::min / std::max NaN semantics
            v = std_max(v, dg.cmin[ch]);
        }
        d.U_shift[e] = v;
        d.U[e] = v;        // publication: m_optimal_control = m_optimal_control_shifted (mppi.cpp:178-182)
        d.U_snap[e] = v;   // input of the optimal re-rollout (side stream)
    }
    // Publication to the host-mapped block: the LAST channel block to get here copies the whole sequence with
    // 16-byte stores in address order (full PCIe write segments) instead of nu x T strided 8-byte stores.
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(d.finish_count, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const double2 *src = reinterpret_cast<const double2 *>(d.U);
        double2 *dst = reinterpret_cast<double2 *>(d.result);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) dst[i] = __ldcg(src + i);
        if ((n & 1) && threadIdx.x == 0) d.result[n - 1] = __ldcg(d.U + n - 1);
        if (threadIdx.x == 0) { finish_publish_stats(d, px, n, staged ? s_comb : nullptr, s_arg); *d.finish_count = 0; }
    }
}

// ---- FMA-chain microbenchmark: the roofline denominator of the rollout kernel -------------------------
template <class R> __global__ void __launch_bounds__(256) k_fma_peak(R *out, int iters) {
    R a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = R(threadIdx.x + i) * R(1e-3);
    const R m = R(0.999), c = R(1e-6);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * m + c;
    }
    R s = R(0);
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_fma_peak(int precision, double *tflops) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    void *buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)blocks * threads * 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        if (precision == 0) k_fma_peak<double><<<blocks, threads>>>((double *)buf, iters); else k_fma_peak<float><<<blocks, threads>>>((float *)buf, iters);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return cudaGetLastError();
}

// ---- launchers -----------------------------------------------------------------------------------------
cudaError_t launch_select_kept(const DeviceState &d, cudaStream_t s) {
    return launch_chain(k_select_kept, dim3(1, d.batch), dim3(1024), 0, s, d);
}
cudaError_t launch_merge_kept(const DeviceState &d, cudaStream_t s) {
    return launch_chain(k_merge_kept, dim3(1), dim3(256), 0, s, d);
}

template <class R, class RI, int NU> static cudaError_t sample_tt(const DeviceState &d, cudaStream_t s, int *launches) {
    const size_t row = sizeof(R) * (size_t)d.nu * d.T;
    if (d.keep_best > 0) {
        if (row > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_shift_kept<R, RI, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row);
            if (e != cudaSuccess) return e;
        }
        { cudaError_t e = launch_chain(k_shift_kept<R, RI, NU>, dim3((unsigned)d.keep_best, d.batch), dim3(64), row, s, d); if (e != cudaSuccess) return e; }
        ++*launches;
    }
    if (d.chase) return cudaSuccess;   // the columns come from the sampling warps of the rollout blocks (sample_core.cuh: chase_sampler)
    const long long ncols = d.k_count * d.T;
    if constexpr (NU % 4 == 0) {
        // MPPI_B200_SAMPLE_TILE=1 keeps the general kernel, =2 the one-thread-per-Philox-block kernel (A/B measurements)
        static const int sample_switch = std::getenv("MPPI_B200_SAMPLE_TILE") ? std::atoi(std::getenv("MPPI_B200_SAMPLE_TILE")) : 0;
        const bool tile_only = sample_switch == 1;
        if (d.L_is_diagonal && sample_switch == 0) {
            ++*launches;
            return launch_chain(k_sample_columns<R, RI, NU>, dim3((unsigned)((ncols + 255) / 256) + 1, d.batch), dim3(256), 0, s, d);   // + the prepare block
        }
        if (d.L_is_diagonal && !tile_only) {
            const long long quads = ncols * (NU / 4);
            ++*launches;
            return launch_chain(k_sample_quads<R, RI, NU>, dim3((unsigned)((quads + 255) / 256) + 1, d.batch), dim3(256), 0, s, d);   // + the prepare block
        }
    }
    const unsigned grid = (unsigned)((ncols + 255) / 256);
    const size_t tile = sizeof(R) * 256 * (size_t)NU;
    ++*launches;
    return launch_chain(k_sample<R, RI, NU>, dim3(grid + 1, d.batch), dim3(256), tile, s, d);   // + the prepare block
}
template <class R> static cudaError_t sample_t(const DeviceState &d, cudaStream_t s, int *launches) {
    if (d.nu == 12) return d.injected_is_double ? sample_tt<R, double, 12>(d, s, launches) : sample_tt<R, R, 12>(d, s, launches);
    if (d.nu == 2) return d.injected_is_double ? sample_tt<R, double, 2>(d, s, launches) : sample_tt<R, R, 2>(d, s, launches);
    return cudaErrorInvalidValue;
}
cudaError_t launch_sample(const DeviceState &d, int precision, cudaStream_t s, int *launches) {
    return precision == 0 ? sample_t<double>(d, s, launches) : sample_t<float>(d, s, launches);
}

cudaError_t launch_rollout(const DeviceState &d, int precision, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query) {
    return precision == 0 ? launch_rollout_f64(d, variant, faithful, params, optimal_only, s, chase_query) : launch_rollout_f32(d, variant, faithful, params, optimal_only, s, chase_query);
}

// ---- peer-memory exchange (kernels.cuh: PeerExchange) ---------------------------------------------------------
// grid = world blocks. Block p != rank: this rank's payload -> slot [rank] of peer p's mailbox, fence, flag.
// Block rank: wait for every peer's flag in the local mailbox, combine the slots in rank order into the exchange buffer
// (MAX for {-min, max, valid}, SUM for {sum w, sum w*eps, argmin slots}; the warm-start candidates are concatenated).
__global__ void __launch_bounds__(256) k_exchange(const __grid_constant__ DeviceState d, const __grid_constant__ PeerExchange px, int kind) {
    pdl_wait();
    const int count = px.count[kind];
    const unsigned long long update = d.frame->attempt;
    const int parity = (int)(update & 1ull);
    const unsigned long long seq = update * 4ull + (unsigned long long)kind + 1ull;   // never 0, unique per (update, kind)
    double *payload = kind == EX_MINMAX ? d.minmax_local : (kind == EX_SUMS ? d.sums : d.cand);
    const long long slot0 = px.offset[parity][kind];
    const long long flag0 = px.flags_offset + ((long long)parity * EX_KINDS + kind) * px.world;
    const int p = blockIdx.x;
    if (p != px.rank) {
        double *dst = px.mail[p] + slot0 + (long long)px.rank * count;
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = payload[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(px.mail[p] + flag0) + px.rank;
            *flag = seq;
            atomicAdd(px.copies_done, 1);   // the payload has been read: the combining block may overwrite it
        }
        return;
    }
    const double *mine = px.mail[px.rank] + slot0;
    if (threadIdx.x < px.world && threadIdx.x != px.rank) {
        const volatile unsigned long long *flag = reinterpret_cast<const volatile unsigned long long *>(px.mail[px.rank] + flag0) + threadIdx.x;
        const long long t0 = clock64();
        while (*flag != seq) {
            if (*px.error_dev || clock64() - t0 > px.timeout_cycles) { *px.error = 1; *px.error_dev = 1; break; }   // never hang the device on a missing peer
            __nanosleep(100);
        }
    }
    if (threadIdx.x == 0) {   // the (co-resident) copy blocks of this launch are done with the payload
        const long long t0 = clock64();
        while (atomicAdd(px.copies_done, 0) != px.world - 1) { if (clock64() - t0 > px.timeout_cycles) { *px.error = 1; *px.error_dev = 1; break; } }
        *px.copies_done = 0;
    }
    __threadfence_system();
    __syncthreads();
    if (kind == EX_CAND) {
        for (int i = threadIdx.x; i < count * px.world; i += blockDim.x) {
            const int q = i / count, e = i - q * count;
            d.cand_all[i] = (q == px.rank) ? payload[e] : __ldcg(mine + (long long)q * count + e);
        }
        return;
    }
    for (int e = threadIdx.x; e < count; e += blockDim.x) {
        double acc = (kind == EX_MINMAX) ? -CUDART_INF : 0.0;
        for (int q = 0; q < px.world; q++) {
            const double v = (q == px.rank) ? payload[e] : __ldcg(mine + (long long)q * count + e);
            acc = (kind == EX_MINMAX) ? fmax(acc, v) : acc + v;
        }
        payload[e] = acc;
    }
}

cudaError_t launch_exchange(const DeviceState &d, const PeerExchange &px, int kind, cudaStream_t s) {
    return launch_chain(k_exchange, dim3(px.world), dim3(256), 0, s, d, px, kind);
}

cudaError_t launch_weights(const DeviceState &d, cudaStream_t s) {
    return launch_chain(k_weights, dim3(d.weight_blocks, d.batch), dim3(256), 0, s, d);
}

template <class R, bool FUSED> static void (*gradient_kernel(int groups))(DeviceState) {
    return groups == 1 ? k_gradient<R, 1, FUSED> : (groups == 2 ? k_gradient<R, 2, FUSED> : k_gradient<R, 4, FUSED>);
}

cudaError_t launch_gradient(const DeviceState &d, int precision, cudaStream_t s, int *launches) {
    const int n = d.nu * d.T;
    const int nvec = n / (precision == 0 ? 2 : 4);
    int threads = std::min(512, ((nvec + 127) / 128) * 128), groups = 1;
    if (nvec <= 256 && d.k_count >= 4096) { groups = 512 / nvec >= 4 ? 4 : 2; threads = ((groups * nvec + 31) / 32) * 32; }   // short rows: several rows per block pass
    const bool fused = d.fused_tail != 0;
    const size_t smem = groups > 1 ? sizeof(double) * (size_t)(groups - (fused ? 0 : 1)) * n : 0;   // (fused: + the staging row of the channel-major rewrite)
    const dim3 grid(d.grad_blocks, d.batch);
    void (*kern)(DeviceState) = nullptr;
    if (precision == 0) kern = fused ? gradient_kernel<double, true>(groups) : gradient_kernel<double, false>(groups);
    else kern = fused ? gradient_kernel<float, true>(groups) : gradient_kernel<float, false>(groups);
    cudaError_t e = launch_chain(kern, grid, dim3(threads), smem, s, d);
    if (e != cudaSuccess || fused) { *launches += 1; return e; }
    *launches += 2;
    return launch_chain(k_gradient_reduce, dim3((n + 31) / 32, d.batch), dim3(1024), 0, s, d);
}

cudaError_t launch_finish(const DeviceState &d, cudaStream_t s) {
    size_t smem = sizeof(double) * (2 * (size_t)d.T + 2 * (size_t)d.sg_len + 2 * (size_t)d.sg_window + 1);
    if (d.world > 1 && d.has_px) smem += sizeof(double) * (((size_t)d.T + 1) * (size_t)d.world + (size_t)d.world);   // combined + argmin slots + the peers' values of one channel
    if (d.fused_tail) smem += sizeof(double) * (9 * (size_t)d.T + 1);   // combined + up to eight slices of the channel's partial sums
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_chain(k_finish, dim3(d.nu, d.batch), dim3(((d.world > 1 && d.has_px) || d.fused_tail) ? 256 : 64), smem, s, d);
}

}  // namespace mppi_b200
