// Small fixed-size linear algebra for the rollout kernel: everything lives in registers (or
// thread-local memory when ptxas spills), every loop is unrolled at compile time.
// Host/device: the same templates compile for the CPU so tests/ can check them against the
// oracle without a GPU (tests/test_device_math_host.py); the product only ever runs them on
// the device.
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define MPPI_HD __host__ __device__ __forceinline__
#else
#define MPPI_HD inline
#endif

namespace mppi_b200 {

template <class R> struct Vec3 {
    R x, y, z;
};
template <class R> MPPI_HD Vec3<R> v3(R x, R y, R z) { Vec3<R> r; r.x = x; r.y = y; r.z = z; return r; }
template <class R> MPPI_HD Vec3<R> operator+(const Vec3<R> &a, const Vec3<R> &b) { return v3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class R> MPPI_HD Vec3<R> operator-(const Vec3<R> &a, const Vec3<R> &b) { return v3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class R> MPPI_HD Vec3<R> operator*(const Vec3<R> &a, R s) { return v3<R>(a.x * s, a.y * s, a.z * s); }
template <class R> MPPI_HD Vec3<R> cross(const Vec3<R> &a, const Vec3<R> &b) {
    return v3<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class R> MPPI_HD R dot(const Vec3<R> &a, const Vec3<R> &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// row-major 3x3
template <class R> struct Mat3 {
    R m[9];
    MPPI_HD R &operator()(int r, int c) { return m[3 * r + c]; }
    MPPI_HD const R &operator()(int r, int c) const { return m[3 * r + c]; }
};
template <class R> MPPI_HD Vec3<R> mul(const Mat3<R> &a, const Vec3<R> &v) {
    return v3<R>(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
template <class R> MPPI_HD Vec3<R> tmul(const Mat3<R> &a, const Vec3<R> &v) {  // a^T v
    return v3<R>(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
template <class R> MPPI_HD Mat3<R> matmul(const Mat3<R> &a, const Mat3<R> &b) {
    Mat3<R> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
    return r;
}
template <class R> MPPI_HD Mat3<R> matmul_t(const Mat3<R> &a, const Mat3<R> &b) {  // a * b^T
    Mat3<R> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(j, 0) + a(i, 1) * b(j, 1) + a(i, 2) * b(j, 2);
    return r;
}

// symmetric 3x3: xx xy xz yy yz zz
template <class R> struct Sym3 {
    R xx, xy, xz, yy, yz, zz;
};
template <class R> MPPI_HD Vec3<R> mul(const Sym3<R> &s, const Vec3<R> &v) {
    return v3<R>(s.xx * v.x + s.xy * v.y + s.xz * v.z, s.xy * v.x + s.yy * v.y + s.yz * v.z, s.xz * v.x + s.yz * v.y + s.zz * v.z);
}
// R S R^T for symmetric S (result symmetric)
template <class R> MPPI_HD Sym3<R> conj(const Mat3<R> &r, const Sym3<R> &s) {
    // T = R S (3x3), then rows of T dotted with rows of R
    R t[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[3 * i + 0] = r(i, 0) * s.xx + r(i, 1) * s.xy + r(i, 2) * s.xz;
        t[3 * i + 1] = r(i, 0) * s.xy + r(i, 1) * s.yy + r(i, 2) * s.yz;
        t[3 * i + 2] = r(i, 0) * s.xz + r(i, 1) * s.yz + r(i, 2) * s.zz;
    }
    Sym3<R> o;
    o.xx = t[0] * r(0, 0) + t[1] * r(0, 1) + t[2] * r(0, 2);
    o.xy = t[0] * r(1, 0) + t[1] * r(1, 1) + t[2] * r(1, 2);
    o.xz = t[0] * r(2, 0) + t[1] * r(2, 1) + t[2] * r(2, 2);
    o.yy = t[3] * r(1, 0) + t[4] * r(1, 1) + t[5] * r(1, 2);
    o.yz = t[3] * r(2, 0) + t[4] * r(2, 1) + t[5] * r(2, 2);
    o.zz = t[6] * r(2, 0) + t[7] * r(2, 1) + t[8] * r(2, 2);
    return o;
}
template <class R> MPPI_HD Mat3<R> conj(const Mat3<R> &r, const Mat3<R> &b) { return matmul_t(matmul(r, b), r); }

// spatial motion [linear; angular] and force [linear; angular]
template <class R> struct Mot {
    Vec3<R> v, w;
};
template <class R> struct Frc {
    Vec3<R> f, n;
};

// rigid transform child -> parent
template <class R> struct Xf {
    Mat3<R> R_;
    Vec3<R> p;
};
template <class R> MPPI_HD Mot<R> act(const Xf<R> &M, const Mot<R> &m) { Mot<R> o; o.w = mul(M.R_, m.w); o.v = mul(M.R_, m.v) + cross(M.p, o.w); return o; }
template <class R> MPPI_HD Mot<R> act_inv(const Xf<R> &M, const Mot<R> &m) { Mot<R> o; o.v = tmul(M.R_, m.v - cross(M.p, m.w)); o.w = tmul(M.R_, m.w); return o; }
template <class R> MPPI_HD Frc<R> act(const Xf<R> &M, const Frc<R> &f) { Frc<R> o; o.f = mul(M.R_, f.f); o.n = mul(M.R_, f.n) + cross(M.p, o.f); return o; }
template <class R> MPPI_HD Mot<R> mcross(const Mot<R> &a, const Mot<R> &b) { Mot<R> o; o.v = cross(a.w, b.v) + cross(a.v, b.w); o.w = cross(a.w, b.w); return o; }
template <class R> MPPI_HD Frc<R> fcross(const Mot<R> &a, const Frc<R> &f) { Frc<R> o; o.f = cross(a.w, f.f); o.n = cross(a.w, f.n) + cross(a.v, f.f); return o; }

// articulated inertia: f = A v + B w, n = B^T v + D w  (A, D symmetric)
template <class R> struct Art {
    Sym3<R> A, D;
    Mat3<R> B;
};

// precision-generic math
MPPI_HD void sincos_(double a, double *s, double *c) {
#if defined(__CUDA_ARCH__)
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#endif
}
MPPI_HD void sincos_(float a, float *s, float *c) {
#if defined(__CUDA_ARCH__)
    sincosf(a, s, c);
#else
    *s = sinf(a); *c = cosf(a);
#endif
}
// fused multiply-add, spelled out where the operation order matters for the instruction count: the compiler contracts
// a * b + c on the device but may not re-associate  x - (a * b - c * d)  into two dependent FMAs
MPPI_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
MPPI_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
// the word that carries the sign bit (bit 31)
MPPI_HD unsigned int sign_word(double a) {
#if defined(__CUDA_ARCH__)
    return (unsigned int)__double2hiint(a);
#else
    unsigned long long b; memcpy(&b, &a, sizeof b); return (unsigned int)(b >> 32);
#endif
}
MPPI_HD unsigned int sign_word(float a) {
#if defined(__CUDA_ARCH__)
    return (unsigned int)__float_as_int(a);
#else
    unsigned int b; memcpy(&b, &a, sizeof b); return b;
#endif
}
// x < 0 decided on the integer pipe: sign bit set and not a NaN (-inf and -0 included; -0 never arises from a difference
// of finite numbers). Same truth value as the FP comparison for every non-zero x, NaN included (false).
MPPI_HD bool is_negative(double x) { return (sign_word(x) ^ 0x80000000u) <= 0x7ff00000u; }
MPPI_HD bool is_negative(float x) { return (sign_word(x) ^ 0x80000000u) <= 0x7f800000u; }
MPPI_HD int float_bits(float a) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(a);
#else
    int b; memcpy(&b, &a, sizeof b); return b;
#endif
}
MPPI_HD float xor_high(float a, unsigned int mask) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(__float_as_int(a) ^ (int)mask);
#else
    unsigned int b; memcpy(&b, &a, sizeof b); b ^= mask; memcpy(&a, &b, sizeof b); return a;
#endif
}
// low 32 bits of a double; v with the bits of `mask` (bit 31 = sign) flipped in its high word
MPPI_HD int low_word(double a) {
#if defined(__CUDA_ARCH__)
    return __double2loint(a);
#else
    unsigned long long b; memcpy(&b, &a, sizeof b); return (int)(unsigned int)(b & 0xffffffffull);
#endif
}
MPPI_HD double xor_high(double a, unsigned int mask) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(a) ^ (int)mask, __double2loint(a));
#else
    unsigned long long b; memcpy(&b, &a, sizeof b); b ^= (unsigned long long)mask << 32; memcpy(&a, &b, sizeof b); return a;
#endif
}
MPPI_HD double sqrt_(double a) { return sqrt(a); }
MPPI_HD float sqrt_(float a) { return sqrtf(a); }
MPPI_HD double acos_(double a) { return acos(a); }
MPPI_HD float acos_(float a) { return acosf(a); }
MPPI_HD double exp_(double a) { return exp(a); }
MPPI_HD float exp_(float a) { return expf(a); }
MPPI_HD double log10_(double a) { return log10(a); }
MPPI_HD float log10_(float a) { return log10f(a); }
MPPI_HD double fabs_(double a) { return fabs(a); }
MPPI_HD float fabs_(float a) { return fabsf(a); }
MPPI_HD double fmin_(double a, double b) { return a < b ? a : b; }   // std::min(a,b): b<a ? b : a
MPPI_HD float fmin_(float a, float b) { return a < b ? a : b; }
MPPI_HD double copysign_(double a, double b) { return copysign(a, b); }
MPPI_HD float copysign_(float a, float b) { return copysignf(a, b); }
// std::min / std::max with the reference's argument order semantics (NaN handling preserved)
template <class R> MPPI_HD R std_min(R a, R b) { return (b < a) ? b : a; }
template <class R> MPPI_HD R std_max(R a, R b) { return (a < b) ? b : a; }
template <class R> MPPI_HD R std_clamp(R v, R lo, R hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

// a x b + c, each component one chain of two fused multiply-adds (the expression form is multiply, FMA, add)
template <class R> MPPI_HD Vec3<R> cross_add(const Vec3<R> &a, const Vec3<R> &b, const Vec3<R> &c) {
    return v3<R>(fma_(a.y, b.z, fma_(-a.z, b.y, c.x)), fma_(a.z, b.x, fma_(-a.x, b.z, c.y)), fma_(a.x, b.y, fma_(-a.y, b.x, c.z)));
}
// c - a x b
template <class R> MPPI_HD Vec3<R> cross_sub(const Vec3<R> &a, const Vec3<R> &b, const Vec3<R> &c) {
    return v3<R>(fma_(a.z, b.y, fma_(-a.y, b.z, c.x)), fma_(a.x, b.z, fma_(-a.z, b.x, c.y)), fma_(a.y, b.x, fma_(-a.x, b.y, c.z)));
}

// The same two with structural zeros in `a` (bit k of MASK clear = component k is zero by the robot's structure): the
// products with those components are dropped. The compiler may not do that by itself (0 * inf is NaN, -0 + 0 is +0), so
// a literal or constant-memory zero still costs its multiply-add — and, in a dependency chain, its latency.
template <bool NZ, class R> MPPI_HD R fma_nz(R a, R b, R c) { if (NZ) return fma_(a, b, c); return c; }
template <unsigned MASK, class R> MPPI_HD Vec3<R> cross_add_m(const Vec3<R> &a, const Vec3<R> &b, const Vec3<R> &c) {
    constexpr bool X = (MASK & 1u) != 0, Y = (MASK & 2u) != 0, Z = (MASK & 4u) != 0;
    return v3<R>(fma_nz<Y>(a.y, b.z, fma_nz<Z>(-a.z, b.y, c.x)), fma_nz<Z>(a.z, b.x, fma_nz<X>(-a.x, b.z, c.y)), fma_nz<X>(a.x, b.y, fma_nz<Y>(-a.y, b.x, c.z)));
}
template <unsigned MASK, class R> MPPI_HD Vec3<R> cross_sub_m(const Vec3<R> &a, const Vec3<R> &b, const Vec3<R> &c) {
    constexpr bool X = (MASK & 1u) != 0, Y = (MASK & 2u) != 0, Z = (MASK & 4u) != 0;
    return v3<R>(fma_nz<Z>(a.z, b.y, fma_nz<Y>(-a.y, b.z, c.x)), fma_nz<X>(a.x, b.z, fma_nz<Z>(-a.z, b.x, c.y)), fma_nz<Y>(a.y, b.x, fma_nz<X>(-a.x, b.y, c.z)));
}

// ---- plane rotations ------------------------------------------------------------------------------
template <class R> MPPI_HD Vec3<R> rotz(R c, R s, const Vec3<R> &v) { return v3<R>(c * v.x - s * v.y, s * v.x + c * v.y, v.z); }
template <class R> MPPI_HD Vec3<R> rotz_t(R c, R s, const Vec3<R> &v) { return v3<R>(c * v.x + s * v.y, c * v.y - s * v.x, v.z); }
template <class R> MPPI_HD Vec3<R> rotx(R c, R s, const Vec3<R> &v) { return v3<R>(v.x, c * v.y - s * v.z, s * v.y + c * v.z); }
template <class R> MPPI_HD Vec3<R> rotx_t(R c, R s, const Vec3<R> &v) { return v3<R>(v.x, c * v.y + s * v.z, c * v.z - s * v.y); }


}  // namespace mppi_b200
