// K1 sample — what one thread produces (reference src/controller/mppi.cpp:222-269 for the two static rollouts and
// the kept set, controller/gaussian.hpp:70-75 for a fresh column eps = V sqrt(Lambda) z). Host/device templates: the
// kernels of k_misc.cu wrap them, and tests/host_check compiles them for the CPU to check the index arithmetic
// (which rollout, which step, which counter, which element) without a GPU. Device arithmetic is the MUFU-based
// single-precision Box–Muller; the host stand-ins (libm) exist for that index check only.
#pragma once
#include <math.h>

#include "kernels.cuh"

namespace mppi_b200 {

MPPI_HD unsigned mulhi_u32(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (unsigned)(((unsigned long long)a * (unsigned long long)b) >> 32);
#endif
}

// ---- Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (column lo, column hi, block, update) --
MPPI_HD uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        // one 32 x 32 -> 64 multiply per product (IMAD.WIDE) instead of a high and a low one
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * ctr.x, p1 = (unsigned long long)0xCD9E8D57u * ctr.z;
        const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}

// two uniforms -> two standard normals (Box–Muller, single precision arithmetic; noise only needs
// to be N(0,1) and reproducible, and the exact values are read back for any parity check)
MPPI_HD void box_muller(unsigned a, unsigned b, float *z0, float *z1) {
    const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0,1]
    const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    float s, c;
#if defined(__CUDA_ARCH__)
    // MUFU-based log / sqrt / sin / cos: the angle 2*pi*u2 - pi stays in [-pi, pi] where __sincosf is accurate to 2^-21
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));   // one MUFU (2 ulp) instead of the IEEE square root's refinement
    __sincosf(6.283185307179586f * u2 - 3.14159265358979f, &s, &c);
#else
    const float r = sqrtf(-2.0f * logf(u1));
    const float a2 = 6.283185307179586f * u2 - 3.14159265358979f;
    s = sinf(a2); c = cosf(a2);
#endif
    *z0 = r * c; *z1 = r * s;
}

// The counter of Philox block `b` of column (global rollout kg, step t): the stream depends on the GLOBAL rollout index
// only, so a sharded rollout set draws the same noise whatever the number of ranks.
MPPI_HD void philox_quad(const DeviceState &d, long long kg, int t, int b, float z[4]) {
    const unsigned long long col = (unsigned long long)kg * (unsigned long long)d.T + (unsigned long long)t;
    const uint2 key = make_uint2((unsigned)d.frame->seed, (unsigned)(d.frame->seed >> 32));
    const unsigned upd = (unsigned)d.frame->update_index;
    const uint4 r = philox4x32_10(make_uint4((unsigned)col, (unsigned)(col >> 32), (unsigned)b, upd), key);
    box_muller(r.x, r.y, &z[0], &z[1]);
    box_muller(r.z, r.w, &z[2], &z[3]);
}

// eps_i = sqrt(Sigma_ii) z_i in the arithmetic of the noise buffer: FP64 buffers multiply in FP64; FP32 buffers multiply
// in FP32 (one FMUL instead of two conversions and a DMUL per value — the sampling kernel is bound by its instruction
// count). Every sampling path (column kernel, tile kernel, the kept rollouts' resampled tails) goes through here.
MPPI_HD double scale_noise(double l, float z, double *) { return l * (double)z; }
MPPI_HD float scale_noise(double l, float z, float *) { return (float)l * z; }

// Diagonal transform (every reference configuration, base.hpp:79-83) and NU a multiple of four: elements
// [4b, 4b+4) of column (local rollout kl, step t) depend on ONE Philox block, so one thread produces them and the
// threads of a warp write 32 consecutive quads. Returns false when the values in place stay: a kept rollout, whose
// surviving columns k_shift_kept has already moved and whose tail it has already resampled (mppi.cpp:243-252).
// `ldiag` = the engine's Ldiag (the kernel passes the copy in its parameter bank: the index depends on the thread).
template <class R, class RI, int NU> MPPI_HD bool sample_quad(const DeviceState &d, const double *ldiag, long long kl, int t, int b, R v[4]) {
    static_assert(NU % 4 == 0, "one Philox block per four channels");
    const long long kg = kl + d.k_begin;
    if (kg == 0) {                 // rollout 0: zero noise, always (mppi.cpp:222)
        v[0] = v[1] = v[2] = v[3] = R(0);
        return true;
    }
    if (kg == 1) {                 // rollout 1 = -U_prev, the UNSHIFTED optimum (mppi.cpp:269)
        const double *u = d.U + t * NU + 4 * b;
        v[0] = (R)(-u[0]); v[1] = (R)(-u[1]); v[2] = (R)(-u[2]); v[3] = (R)(-u[3]);
        return true;
    }
    if (d.kept[kl]) return false;
    if (d.frame->noise_source != 0) {
        const RI *src = static_cast<const RI *>(d.injected) + ((size_t)kl * d.T + t) * NU + 4 * b;
        v[0] = (R)src[0]; v[1] = (R)src[1]; v[2] = (R)src[2]; v[3] = (R)src[3];
        return true;
    }
    float z[4];
    philox_quad(d, kg, t, b, z);
    const double *l = ldiag + 4 * b;
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = scale_noise(l[i], z[i], (R *)nullptr);
    return true;
}

// The same values for a whole column (local rollout kl, step t) by ONE thread: the case analysis, the kept flag, the
// frame reads and the index arithmetic are paid once per NU values instead of once per four (k_sample_columns).
// ldiag: constant indices only, so the engine's copy stays in the kernel's parameter bank.
template <class R, class RI, int NU> MPPI_HD bool sample_column(const DeviceState &d, const double *ldiag, long long kl, int t, R v[NU]) {
    static_assert(NU % 4 == 0, "one Philox block per four channels");
    const long long kg = kl + d.k_begin;
    if (kg == 0) {
#pragma unroll
        for (int i = 0; i < NU; i++) v[i] = R(0);
        return true;
    }
    if (kg == 1) {
        const double *u = d.U + t * NU;
#pragma unroll
        for (int i = 0; i < NU; i++) v[i] = (R)(-u[i]);
        return true;
    }
    if (d.kept[kl]) return false;
    if (d.frame->noise_source != 0) {
        const RI *src = static_cast<const RI *>(d.injected) + ((size_t)kl * d.T + t) * NU;
#pragma unroll
        for (int i = 0; i < NU; i++) v[i] = (R)src[i];
        return true;
    }
    const unsigned long long col = (unsigned long long)kg * (unsigned long long)d.T + (unsigned long long)t;
    const uint2 key = make_uint2((unsigned)d.frame->seed, (unsigned)(d.frame->seed >> 32));
    const unsigned upd = (unsigned)d.frame->update_index;
#pragma unroll
    for (int b = 0; b < NU / 4; b++) {
        const uint4 r = philox4x32_10(make_uint4((unsigned)col, (unsigned)(col >> 32), (unsigned)b, upd), key);
        float z[4];
        box_muller(r.x, r.y, &z[0], &z[1]);
        box_muller(r.z, r.w, &z[2], &z[3]);
#pragma unroll
        for (int i = 0; i < 4; i++) v[4 * b + i] = scale_noise(ldiag[4 * b + i], z[i], (R *)nullptr);
    }
    return true;
}

// column index g of a [k_count][T] enumeration -> (local rollout, step); a 32-bit division whenever the enumeration fits
MPPI_HD void column_coordinates(long long g, long long cols, int T, long long *kl, int *t) {
    if (cols <= 0x7fffffffLL) {
        const unsigned ug = (unsigned)g, k = ug / (unsigned)T;
        *t = (int)(ug - k * (unsigned)T); *kl = (long long)k;
    } else {
        const long long k = g / T;
        *t = (int)(g - k * T); *kl = k;
    }
}

// quad index g of a [k_count][T][NU/4] enumeration -> (local rollout, step, Philox block); 32-bit divisions
// whenever the enumeration fits (a 64-bit division is ~100 instructions on the device)
template <int NU> MPPI_HD void quad_coordinates(long long g, long long quads, int T, long long *kl, int *t, int *b) {
    constexpr int NB = NU / 4;
    if (quads <= 0x7fffffffLL) {
        const unsigned ug = (unsigned)g, col = ug / (unsigned)NB, k = col / (unsigned)T;
        *b = (int)(ug - col * NB); *t = (int)(col - k * (unsigned)T); *kl = (long long)k;
    } else {
        const long long col = g / NB, k = col / T;
        *b = (int)(g - col * NB); *t = (int)(col - k * T); *kl = k;
    }
}


// Column i of a rollout block's chunk-major enumeration (noise chase, rollout_core.cuh): chunk c = steps [c CHASE_STEPS,
// (c + 1) CHASE_STEPS) of the block's 32 rollouts; within a chunk consecutive columns are consecutive steps of one rollout.
MPPI_HD void chase_column_coordinates(int i, int *chunk, int *rollout, int *t) {
    const int c = i / CHASE_COLUMNS, within = i - c * CHASE_COLUMNS;
    *chunk = c; *rollout = within / CHASE_STEPS; *t = c * CHASE_STEPS + within % CHASE_STEPS;
}

#if defined(__CUDACC__)
// One column's NU values to the noise buffer. 32-byte stores (sm_100: STG.256) wherever the address allows: every store
// then fills whole 32-byte sectors. With 16-byte stores the column's sectors arrived in halves from different
// instructions and the write stream stalled at 3.5 TB/s whatever the instruction count. `odd`: parity of the column's
// index in the buffer (FP32 columns are 48 bytes: even ones start on a sector, odd ones 16 bytes into one).
template <class R, int NU> __device__ __forceinline__ void store_column(R *dst, const R *v, bool odd) {
    if constexpr (sizeof(R) == 8 && NU % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NU / 4; i++)   // NU * 8 bytes per column: a multiple of 32, so every column starts on a sector
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "d"((double)v[4 * i]), "d"((double)v[4 * i + 1]), "d"((double)v[4 * i + 2]), "d"((double)v[4 * i + 3]) : "memory");
    } else if constexpr (sizeof(R) == 4 && NU == 12) {
        const float *f = reinterpret_cast<const float *>(v);
        if (!odd) {
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
            *reinterpret_cast<float4 *>(dst + 8) = make_float4(f[8], f[9], f[10], f[11]);
        } else {
            *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]);
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 4), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]), "f"(f[8]), "f"(f[9]), "f"(f[10]), "f"(f[11]) : "memory");
        }
    } else if constexpr (sizeof(R) == 8) {
#pragma unroll
        for (int i = 0; i < NU / 2; i++) reinterpret_cast<double2 *>(dst)[i] = make_double2((double)v[2 * i], (double)v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < NU / 4; i++) reinterpret_cast<float4 *>(dst)[i] = make_float4((float)v[4 * i], (float)v[4 * i + 1], (float)v[4 * i + 2], (float)v[4 * i + 3]);
    }
}

// ---- prepare: time shift of the optimal control, rollout 1 = -U_prev, reset of the reductions -------
// mppi.cpp:194-206 (shift), :269 (rollout[1].noise = -m_optimal_control, the UNSHIFTED optimum).
// Runs as the extra last block of k_sample: nothing here is read by the sampling blocks (the two static rollouts
// are produced by the sampling blocks themselves), so one launch and one dependency level less per update — or, with the
// noise chase, on the sampling warps of the rollout grid's first block (threads tid of nthreads).
__device__ __forceinline__ void prepare_block(const DeviceState &d, int tid, int nthreads) {
    const int n = d.nu * d.T;
    const long long shift = d.frame->shift_by;
    for (int i = tid; i < d.frame_doubles; i += nthreads) d.frame_snap[i] = reinterpret_cast<const double *>(d.frame)[i];
    if (shift > 0) {
        for (int e = tid; e < n; e += nthreads) {
            const int t = e / d.nu, dd = e - t * d.nu;
            const long long shifted = d.T - shift;  // columns that survive
            const int src_t = (t < shifted) ? (int)(t + shift) : d.T - 1;
            d.U_shift[e] = d.U[src_t * d.nu + dd];
        }
    }
    if (tid == 0) {
        d.minmax_enc[0] = 0xffffffffffffffffull;
        d.minmax_enc[1] = 0ull;
        *d.valid_count = 0;
        *d.argmin = 0x7fffffffffffffffll;
        *d.skip = 0;
    }
}


// ---- sampling warps of a rollout block (noise chase, rollout_core.cuh) ------------------------------------------------
// Thread j of the `ns` sampling threads of the block that integrates rollouts [first, first + 32). The block's columns are
// enumerated chunk-major — chunk c = steps [c CHASE_STEPS, (c+1) CHASE_STEPS) of the 32 rollouts, CHASE_COLUMNS columns — and
// thread j takes columns j, j + ns, ...: the first chunks are complete after one column time. Every column, drawn or not
// (kept rollout, past the end of the set or of the horizon), is counted in its chunk's shared-memory counter after a
// block-scope fence; the rollout warp waits for CHASE_COLUMNS (noise_chase_wait). Same counters, same arithmetic as
// k_sample_columns: identical noise.
template <class R, int NU> __device__ __forceinline__ void chase_sampler(const DeviceState &d, const double *ldiag, long long first, int j, int ns) {
    const int chunks = (d.T + CHASE_STEPS - 1) / CHASE_STEPS;
    for (int i = j; i < chunks * CHASE_COLUMNS; i += ns) {
        int c, r, t;
        chase_column_coordinates(i, &c, &r, &t);
        const long long kl = first + r;
        if (kl < d.k_count && t < d.T) {
            R v[NU];
            const bool fresh = d.injected_is_double ? sample_column<R, double, NU>(d, ldiag, kl, t, v) : sample_column<R, R, NU>(d, ldiag, kl, t, v);
            if (fresh) {
                const long long col = kl * d.T + t;
                store_column<R, NU>(static_cast<R *>(d.noise) + (size_t)col * NU, v, (col & 1) != 0);
            }
        }
        __threadfence_block();
        atomicAdd(s_chase_columns + c, 1u);
    }
}
#endif

}  // namespace mppi_b200
