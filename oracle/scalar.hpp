// TEST INFRASTRUCTURE — CPU oracle for the MPPI rollout path. Not part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, link or call anything under oracle/.
//
// Scalar abstraction: the oracle is templated on its scalar so the same code runs as plain
// FP64 (the parity reference) and as an op-counting scalar ("instrument the
// oracle with an op-counting scalar once" to fix the algorithmic flops per rollout-step).
#pragma once
#include <cmath>
#include <cstdint>
#include <algorithm>

namespace oracle {

struct OpCount {
    static inline thread_local std::uint64_t addsub = 0, mul = 0, div = 0, sqrt_ = 0, trans = 0, cmp = 0;
    static void reset() { addsub = mul = div = sqrt_ = trans = cmp = 0; }
    // Convention of SURVEY §8d: add/sub/mul/div/sqrt count 1 each, a transcendental counts 1.
    static std::uint64_t flops() { return addsub + mul + div + sqrt_ + trans; }
};

struct Counted {
    double x;
    Counted() : x(0.0) {}
    Counted(double v) : x(v) {}
    explicit operator double() const { return x; }
};
inline Counted operator+(Counted a, Counted b) { ++OpCount::addsub; return Counted(a.x + b.x); }
inline Counted operator-(Counted a, Counted b) { ++OpCount::addsub; return Counted(a.x - b.x); }
inline Counted operator*(Counted a, Counted b) { ++OpCount::mul; return Counted(a.x * b.x); }
inline Counted operator/(Counted a, Counted b) { ++OpCount::div; return Counted(a.x / b.x); }
inline Counted operator-(Counted a) { return Counted(-a.x); }
inline Counted &operator+=(Counted &a, Counted b) { a = a + b; return a; }
inline Counted &operator-=(Counted &a, Counted b) { a = a - b; return a; }
inline Counted &operator*=(Counted &a, Counted b) { a = a * b; return a; }
inline bool operator<(Counted a, Counted b) { ++OpCount::cmp; return a.x < b.x; }
inline bool operator>(Counted a, Counted b) { ++OpCount::cmp; return a.x > b.x; }
inline bool operator<=(Counted a, Counted b) { ++OpCount::cmp; return a.x <= b.x; }
inline bool operator>=(Counted a, Counted b) { ++OpCount::cmp; return a.x >= b.x; }

// math wrappers (overloaded for double and Counted)
inline double m_val(double a) { return a; }
inline double m_val(Counted a) { return a.x; }
inline double m_sqrt(double a) { return std::sqrt(a); }
inline Counted m_sqrt(Counted a) { ++OpCount::sqrt_; return Counted(std::sqrt(a.x)); }
inline double m_sin(double a) { return std::sin(a); }
inline Counted m_sin(Counted a) { ++OpCount::trans; return Counted(std::sin(a.x)); }
inline double m_cos(double a) { return std::cos(a); }
inline Counted m_cos(Counted a) { ++OpCount::trans; return Counted(std::cos(a.x)); }
inline double m_acos(double a) { return std::acos(a); }
inline Counted m_acos(Counted a) { ++OpCount::trans; return Counted(std::acos(a.x)); }
inline double m_exp(double a) { return std::exp(a); }
inline Counted m_exp(Counted a) { ++OpCount::trans; return Counted(std::exp(a.x)); }
inline double m_log10(double a) { return std::log10(a); }
inline Counted m_log10(Counted a) { ++OpCount::trans; return Counted(std::log10(a.x)); }
inline double m_pow(double a, double b) { return std::pow(a, b); }
inline Counted m_pow(Counted a, double b) { ++OpCount::mul; return Counted(std::pow(a.x, b)); }
inline double m_fabs(double a) { return std::fabs(a); }
inline Counted m_fabs(Counted a) { return Counted(std::fabs(a.x)); }
inline double m_min(double a, double b) { return std::min(a, b); }
inline Counted m_min(Counted a, Counted b) { ++OpCount::cmp; return Counted(std::min(a.x, b.x)); }
inline double m_max(double a, double b) { return std::max(a, b); }
inline Counted m_max(Counted a, Counted b) { ++OpCount::cmp; return Counted(std::max(a.x, b.x)); }
inline bool m_isnan(double a) { return std::isnan(a); }
inline bool m_isnan(Counted a) { return std::isnan(a.x); }
inline double m_copysign(double a, double b) { return std::copysign(a, b); }
inline Counted m_copysign(Counted a, Counted b) { return Counted(std::copysign(a.x, b.x)); }

}  // namespace oracle
