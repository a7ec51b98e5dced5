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

}  // namespace mppi_b200
