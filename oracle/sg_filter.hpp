// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header).
// Restatement of the Savitzky–Golay smoothing used by mppi::Trajectory::optimise:
//   gram polynomial weights  — reference src/controller/gram_savitzky_golay/gram_savitzky_golay.cpp:12-53
//   convolution              — gram_savitzky_golay.h:137-152
//   stateful sliding window  — src/controller/filter.cpp:19-116
//   per-channel filter       — src/controller/filter.cpp:118-173
// All quirks of SURVEY Appendix A-5 are kept (write-back one slot early, size_t rotate offset,
// exact double compares). Pinned against the reference's own sources compiled in oracle/_ref.
#pragma once
#include <algorithm>
#include <cstddef>
#include <sstream>
#include <stdexcept>
#include <vector>
#include <cmath>

namespace oracle {

inline double sg_gram_poly(int i, int m, int k, int s) {
    if (k > 0) {
        return (4. * k - 2.) / (k * (2. * m - k + 1.)) * (i * sg_gram_poly(i, m, k - 1, s) + s * sg_gram_poly(i, m, k - 1, s - 1)) -
               ((k - 1.) * (2. * m + k)) / (k * (2. * m - k + 1.)) * sg_gram_poly(i, m, k - 2, s);
    }
    return (k == 0 && s == 0) ? 1. : 0.;
}

inline double sg_gen_fact(int a, int b) {
    double gf = 1.;
    for (int j = (a - b) + 1; j <= a; j++) gf *= j;
    return gf;
}

inline double sg_weight(int i, int t, int m, int n, int s) {
    double w = 0;
    for (int k = 0; k <= n; ++k)
        w = w + (2 * k + 1) * (sg_gen_fact(2 * m, k) / sg_gen_fact(2 * m + k + 1, k + 1)) * sg_gram_poly(i, m, k, 0) * sg_gram_poly(t, m, k, s);
    return w;
}

inline std::vector<double> sg_compute_weights(int m, int t, int n, int s) {
    std::vector<double> w(2 * (std::size_t)m + 1);
    for (int i = 0; i < 2 * m + 1; ++i) w[(std::size_t)i] = sg_weight(i - m, t, m, n, s);
    return w;
}

// filter.cpp:19-116
struct SgWindow {
    int window;
    double last_trim_t;
    std::size_t start_idx;
    std::vector<double> uu, tt;

    SgWindow(int size, int w) : window(w), last_trim_t(-1), start_idx(w) {
        uu.resize(size + 2 * window + 1, 0);
        tt.resize(size + 2 * window + 1, -1);
    }

    void trim(double t) {
        if (t < last_trim_t) throw std::runtime_error("Resetting the window back in the past.");
        last_trim_t = t;
        std::size_t trim_idx = start_idx;
        for (std::size_t i = 0; i < start_idx; i++) {
            if (tt[i] >= t) { trim_idx = i; break; }
        }
        std::size_t offset = trim_idx - window;
        std::rotate(tt.begin(), tt.begin() + offset, tt.end());
        std::rotate(uu.begin(), uu.begin() + offset, uu.end());
        if (offset > 0) {
            std::fill(tt.end() - offset, tt.end(), *(tt.end() - offset - 1));
            std::fill(uu.end() - offset, uu.end(), *(uu.end() - offset - 1));
        }
        start_idx = window;
        tt[start_idx] = t;
    }

    void add_point(double u, double t) {
        if (t < tt[start_idx]) throw std::runtime_error("Adding measurement older then new time");
        uu[start_idx] = u;
        tt[start_idx] = t;
        std::fill(uu.begin() + start_idx + 1, uu.end(), uu[start_idx]);
        std::fill(tt.begin() + start_idx + 1, tt.end(), tt[start_idx]);
        start_idx++;
    }

    std::size_t lower(double t) const { return std::lower_bound(tt.begin(), tt.end(), t) - tt.begin(); }

    double apply(const std::vector<double> &weights, double t) {
        std::size_t idx = lower(t);
        if (2 * (std::size_t)window + 1 != weights.size()) throw std::logic_error("data to be filtered have wrong size");
        const double *v = uu.data() + idx - window;
        double res = weights[0] * v[0];
        for (std::size_t i = 1; i < weights.size(); ++i) res += weights[i] * v[i];
        res = res / 1.0;  // dt_^derivative_order with s = 0 (gram_savitzky_golay.cpp:61)
        uu[lower(t) - 1] = res;  // filter.cpp:113: writes ONE SLOT EARLIER than the filtered sample
        return res;
    }
};

struct SgFilter {
    std::vector<SgWindow> windows;
    std::vector<double> weights;
    SgFilter(int steps, int nu, int window, unsigned order)
        : windows(nu, SgWindow(steps, window)), weights(sg_compute_weights(window, 0, (int)order, 0)) {}
    void reset(double t) { for (auto &w : windows) w.trim(t); }
    void add_measurement(const double *u, double t) { for (std::size_t i = 0; i < windows.size(); i++) windows[i].add_point(u[i], t); }
    void apply(double *u, double t) { for (std::size_t i = 0; i < windows.size(); i++) u[i] = windows[i].apply(weights, t); }
};

}  // namespace oracle
