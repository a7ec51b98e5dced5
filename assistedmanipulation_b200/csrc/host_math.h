// Host-side arithmetic of engine creation (plain C++: engine.cu includes it, tests/host_check compiles it for the CPU):
// the Savitzky–Golay taps the finish kernel applies and the noise transform the sampling kernels apply.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace mppi_b200 {
namespace host_math {

// gram polynomial Savitzky–Golay weights (Gorry 1990): the taps of gram_sg::SavitzkyGolayFilter(m, t=0, n, s=0)
inline double gram_poly(int i, int m, int k, int s) {
    if (k > 0)
        return (4. * k - 2.) / (k * (2. * m - k + 1.)) * (i * gram_poly(i, m, k - 1, s) + s * gram_poly(i, m, k - 1, s - 1)) -
               ((k - 1.) * (2. * m + k)) / (k * (2. * m - k + 1.)) * gram_poly(i, m, k - 2, s);
    return (k == 0 && s == 0) ? 1. : 0.;
}
inline double gen_fact(int a, int b) { double g = 1.; for (int j = a - b + 1; j <= a; j++) g *= j; return g; }
inline std::vector<double> sg_weights(int m, int n) {
    std::vector<double> w(2 * m + 1);
    for (int i = -m; i <= m; i++) {
        double s = 0;
        for (int k = 0; k <= n; k++) s = s + (2 * k + 1) * (gen_fact(2 * m, k) / gen_fact(2 * m + k + 1, k + 1)) * gram_poly(i, m, k, 0) * gram_poly(0, m, k, 0);
        w[i + m] = s;
    }
    return w;
}

// V * sqrt(Lambda) of the symmetric covariance, eigenvalues ascending (gaussian.hpp:48-55)
inline std::vector<double> noise_transform(int n, const double *cov) {
    std::vector<double> a(cov, cov + (size_t)n * n), V((size_t)n * n, 0.0), ev(n);
    auto A = [&](int r, int c) -> double & { return a[(size_t)c * n + r]; };
    auto E = [&](int r, int c) -> double & { return V[(size_t)c * n + r]; };
    for (int i = 0; i < n; i++) E(i, i) = 1.0;
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A(p, q) * A(p, q);
        if (off == 0.0) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (A(p, q) == 0.0) continue;
                const double theta = (A(q, q) - A(p, p)) / (2.0 * A(p, q));
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) { double x = A(k, p), y = A(k, q); A(k, p) = c * x - s * y; A(k, q) = s * x + c * y; }
                for (int k = 0; k < n; k++) { double x = A(p, k), y = A(q, k); A(p, k) = c * x - s * y; A(q, k) = s * x + c * y; }
                for (int k = 0; k < n; k++) { double x = E(k, p), y = E(k, q); E(k, p) = c * x - s * y; E(k, q) = s * x + c * y; }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = A(i, i);
    for (int i = 0; i < n - 1; i++) {
        int k = 0;
        for (int j = 1; j < n - i; j++) if (ev[i + j] < ev[i + k]) k = j;
        if (k > 0) { std::swap(ev[i], ev[i + k]); for (int r = 0; r < n; r++) std::swap(E(r, i), E(r, i + k)); }
    }
    std::vector<double> L((size_t)n * n);
    for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) L[(size_t)c * n + r] = E(r, c) * std::sqrt(ev[c] > 0 ? ev[c] : 0.0);
    return L;
}


// A DIAGONAL covariance (every reference configuration, base.hpp:79-83) is sampled as eps_i = sqrt(Sigma_ii) z_i. For
// such a matrix the reference's V sqrt(Lambda) (above) is that diagonal with its columns permuted into ascending-eigenvalue
// order: the same distribution with the components of z relabelled. The in-kernel Philox stream has no reference stream
// to follow (the reference-RNG parity mode draws on the host, in the facade, with the reference's own transform), so the
// engine takes the unpermuted form: one multiply per element instead of an nu x nu product per column, and four
// consecutive elements depend on one Philox block (k_sample_quads). Returns false when the matrix is not diagonal.
inline bool diagonal_noise_transform(int n, const double *cov, double *ldiag) {
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++)
            if (r != c && cov[(size_t)c * n + r] != 0.0) return false;
    for (int i = 0; i < n; i++) { const double v = cov[(size_t)i * n + i]; ldiag[i] = std::sqrt(v > 0 ? v : 0.0); }
    return true;
}

}  // namespace host_math
}  // namespace mppi_b200
