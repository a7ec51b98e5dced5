// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header).
// FP64 restatement of mppi::Trajectory (reference src/controller/mppi.cpp:79-512) with the
// rollout indices widened from std::uint8_t to std::size_t (mppi.hpp:639,642; mppi.cpp:222-231,
// 243,255,381,416) — identical behaviour for rollouts + 2 <= 255, and able to run the
// BASELINE.json configurations the unmodified reference cannot.
//
// Noise sources:
//   * GAUSSIAN  — restates controller/gaussian.hpp (std::mt19937 default seed +
//                 std::normal_distribution<double>, transform V*sqrt(L)); stream consumed in the
//                 order of mppi.cpp:243-262. This mode is what oracle/_ref (the reference's own
//                 mppi.cpp compiled against a shim Eigen) is compared with.
//   * INJECTED  — wherever the reference would draw a fresh column m_gaussian() for rollout k,
//                 step t, the column injected[(k*T + t)*nu .. +nu) is used instead. Same
//                 definition as the CUDA engine's injected-noise mode.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <numeric>
#include <optional>
#include <random>
#include <thread>
#include <vector>
#include <stdexcept>
#include <chrono>

#include "sg_filter.hpp"
#include "pool.hpp"

namespace oracle {

// mppi.hpp:30-85
struct Dynamics {
    virtual ~Dynamics() = default;
    virtual std::unique_ptr<Dynamics> copy() = 0;
    virtual const double *step(const double *control, double dt) = 0;
    virtual void set_state(const double *state, double time) = 0;
    virtual const double *get_state() = 0;
    virtual int get_control_dof() = 0;
    virtual int get_state_dof() = 0;
};

// mppi.hpp:93-145
struct Cost {
    virtual ~Cost() = default;
    virtual std::unique_ptr<Cost> copy() = 0;
    virtual void reset(double time) = 0;
    virtual double get_cost(const double *state, const double *control, Dynamics *dynamics, double time) = 0;
    virtual int get_control_dof() = 0;
    virtual int get_state_dof() = 0;
};

// mppi.hpp:181-249
struct Configuration {
    std::vector<double> initial_state;
    std::int64_t rollouts = 1;
    std::int64_t keep_best_rollouts = 0;
    double time_step = 0.01, horison = 1.0, gradient_step = 1.0, cost_scale = 1.0, cost_discount_factor = 1.0;
    std::vector<double> covariance;  // nu x nu, column major
    bool control_bound = false;
    std::vector<double> control_min, control_max;
    std::optional<std::vector<double>> control_default;
    bool smoothing = false;
    unsigned smoothing_window = 10, smoothing_order = 1;
    unsigned threads = 1;
};

// Cyclic Jacobi eigen-solver for the symmetric covariance; eigenvalues sorted ascending with the
// selection-sort-with-swaps that Eigen::SelfAdjointEigenSolver applies (gaussian.hpp:48-55).
// Exact for the diagonal covariances every reference configuration uses (base.hpp:79-83).
inline void symmetric_eigen(int n, std::vector<double> a /* col-major, copied */, std::vector<double> &evals, std::vector<double> &evecs) {
    evecs.assign((std::size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) evecs[(std::size_t)i * n + i] = 1.0;
    auto A = [&](int r, int c) -> double & { return a[(std::size_t)c * n + r]; };
    auto V = [&](int r, int c) -> double & { return evecs[(std::size_t)c * n + r]; };
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A(p, q) * A(p, q);
        if (off == 0.0) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (A(p, q) == 0.0) continue;
                double theta = (A(q, q) - A(p, p)) / (2.0 * A(p, q));
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) { double akp = A(k, p), akq = A(k, q); A(k, p) = c * akp - s * akq; A(k, q) = s * akp + c * akq; }
                for (int k = 0; k < n; k++) { double apk = A(p, k), aqk = A(q, k); A(p, k) = c * apk - s * aqk; A(q, k) = s * apk + c * aqk; }
                for (int k = 0; k < n; k++) { double vkp = V(k, p), vkq = V(k, q); V(k, p) = c * vkp - s * vkq; V(k, q) = s * vkp + c * vkq; }
            }
    }
    evals.resize(n);
    for (int i = 0; i < n; i++) evals[i] = A(i, i);
    for (int i = 0; i < n - 1; i++) {
        int k = 0;
        for (int j = 1; j < n - i; j++) if (evals[i + j] < evals[i + k]) k = j;
        if (k > 0) {
            std::swap(evals[i], evals[i + k]);
            for (int r = 0; r < n; r++) std::swap(V(r, i), V(r, i + k));
        }
    }
}

// controller/gaussian.hpp:12-90
struct Gaussian {
    int n;
    std::vector<double> transform;  // col-major n x n
    std::mt19937 generator;
    std::normal_distribution<double> distribution{0, 1};
    Gaussian(int n_, const std::vector<double> &cov) : n(n_) {
        std::vector<double> ev, V;
        symmetric_eigen(n, cov, ev, V);
        transform.assign((std::size_t)n * n, 0.0);
        for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) transform[(std::size_t)c * n + r] = V[(std::size_t)c * n + r] * std::sqrt(ev[c]);
    }
    void sample(double *out) {
        std::vector<double> z(n);
        for (int i = 0; i < n; i++) z[i] = distribution(generator);
        for (int r = 0; r < n; r++) {
            double s = 0.0;
            for (int c = 0; c < n; c++) s += transform[(std::size_t)c * n + r] * z[c];
            out[r] = 0.0 + s;
        }
    }
};

class Trajectory {
public:
    static constexpr std::int64_t s_static_rollouts = 2;  // mppi.hpp:306

    struct Rollout {
        std::vector<double> noise;  // nu x T column-major: element (d,t) at d + nu*t
        double cost = 0.0;
    };

    // mppi.cpp:11-77 — returns nullptr + reason on the same conditions.
    static std::unique_ptr<Trajectory> create(const Configuration &c, std::unique_ptr<Dynamics> &&dynamics, std::unique_ptr<Cost> &&cost, std::string *why = nullptr) {
        auto fail = [&](const char *m) { if (why) *why = m; return nullptr; };
        if (dynamics->get_control_dof() != cost->get_control_dof()) return fail("controller dynamics control dof != cost control dof");
        if (dynamics->get_state_dof() != cost->get_state_dof()) return fail("controller dynamics state dof != cost state dof");
        int nu = dynamics->get_control_dof();
        if ((int)c.control_min.size() != nu || (int)c.control_max.size() != nu) return fail("controller maximum and minimum must have length control dof");
        if ((int)c.covariance.size() != nu * nu) return fail("controller sample variance dof != dynamics and cost control dof");
        if (c.rollouts < 1) return fail("trajectory rollouts must be greater than zero");
        if (c.keep_best_rollouts < 0) return fail("trajectory cached rollouts cannot be less than zero");
        if (c.threads <= 0) return fail("trajectory threads must be positive nonzero");
        return std::unique_ptr<Trajectory>(new Trajectory(c, std::move(dynamics), std::move(cost)));
    }

    // mppi.cpp:154-187. `injected` == nullptr selects the GAUSSIAN source.
    void update(const double *state, double time, const double *injected = nullptr) {
        using clk = std::chrono::steady_clock;
        m_rollout_state.assign(state, state + m_state_dof);
        m_rollout_time = time;
        auto t0 = clk::now();
        sample(time, injected);
        auto t1 = clk::now();
        rollout();
        auto t2 = clk::now();
        optimise();
        auto t3 = clk::now();
        filter();
        auto t4 = clk::now();
        m_last_rollout_time = m_rollout_time;
        m_optimal_control = m_optimal_control_shifted;
        m_update_duration = std::chrono::duration<double>(clk::now() - t0).count();
        phase_seconds[0] += std::chrono::duration<double>(t1 - t0).count();
        phase_seconds[1] += std::chrono::duration<double>(t2 - t1).count();
        phase_seconds[2] += std::chrono::duration<double>(t3 - t2).count();
        phase_seconds[3] += std::chrono::duration<double>(t4 - t3).count();
        m_update_last = time;
        ++m_update_count;
    }

    // mppi.cpp:481-512
    void get(double *control, double time) const {
        double t = (time - m_last_rollout_time) / m_time_step;
        int lower = (int)t, upper = lower + 1;
        if (upper >= m_step_count) {
            for (int d = 0; d < m_control_dof; d++)
                control[d] = m_control_default ? (*m_control_default)[d] : m_optimal_control[(std::size_t)(m_step_count - 1) * m_control_dof + d];
            return;
        }
        t -= lower;
        for (int d = 0; d < m_control_dof; d++)
            control[d] = (1.0 - t) * m_optimal_control[(std::size_t)lower * m_control_dof + d] + t * m_optimal_control[(std::size_t)upper * m_control_dof + d];
    }

    int step_count() const { return m_step_count; }
    int rollout_count() const { return m_rollout_count; }
    int control_dof() const { return m_control_dof; }
    int state_dof() const { return m_state_dof; }
    const std::vector<Rollout> &rollouts() const { return m_rollouts; }
    const std::vector<double> &weights() const { return m_weights; }
    const std::vector<double> &gradient() const { return m_gradient; }
    const std::vector<double> &optimal() const { return m_optimal_control; }
    // TEST HOOK (no reference counterpart): start the next update from a given published sequence, so that a
    // single-precision engine and this FP64 restatement can be compared update by update from IDENTICAL inputs
    void set_optimal(const double *U) { std::copy(U, U + m_optimal_control.size(), m_optimal_control.begin()); m_optimal_control_shifted = m_optimal_control; }
    double optimal_total_cost() const { return m_optimal_cost; }
    double update_duration() const { return m_update_duration; }
    std::int64_t shift_by() const { return m_shift_by; }
    std::size_t update_count() const { return m_update_count; }   // mppi.hpp:386-388
    Cost &optimal_cost_object() { return *m_cost[0]; }
    double phase_seconds[4] = {0, 0, 0, 0};  // sample / rollout / optimise / filter

private:
    Trajectory(const Configuration &c, std::unique_ptr<Dynamics> &&dynamics, std::unique_ptr<Cost> &&cost)
        : m_step_count((int)std::ceil(c.horison / c.time_step)),
          m_time_step(c.time_step),
          m_rollout_count((int)(c.rollouts + s_static_rollouts)),
          m_thread_count(c.threads),
          m_state_dof(dynamics->get_state_dof()),
          m_control_dof(dynamics->get_control_dof()),
          m_dynamics(c.threads),
          m_cost(c.threads),
          m_gaussian(dynamics->get_control_dof(), c.covariance),
          m_rollout_state(dynamics->get_state_dof(), 0.0),  // initial_state only fixes the size (mppi.cpp:99,121)
          m_cost_discount_factor(c.cost_discount_factor),
          m_cost_scale(c.cost_scale),
          m_weights(m_rollout_count, 0.0),
          m_gradient((std::size_t)m_control_dof * m_step_count, 0.0),
          m_gradient_step(c.gradient_step),
          m_optimal_control_shifted((std::size_t)m_control_dof * m_step_count, 0.0),
          m_optimal_control((std::size_t)m_control_dof * m_step_count, 0.0),
          m_keep_best_rollouts((std::size_t)c.keep_best_rollouts),
          m_ordered_rollouts(c.rollouts),
          m_bound_control(c.control_bound),
          m_control_min(c.control_min),
          m_control_max(c.control_max),
          m_control_default(c.control_default) {
        m_rollouts.resize(m_rollout_count);
        for (auto &r : m_rollouts) r.noise.assign((std::size_t)m_control_dof * m_step_count, 0.0);
        m_dynamics[0] = std::move(dynamics);
        m_cost[0] = std::move(cost);
        for (unsigned i = 1; i < c.threads; i++) {
            m_dynamics[i] = m_dynamics[0]->copy();
            m_cost[i] = m_cost[0]->copy();
        }
        if (c.threads > 1) m_pool = std::make_unique<Pool>(c.threads);
        if (c.smoothing) m_smoothing.emplace(m_step_count, m_control_dof, (int)c.smoothing_window, c.smoothing_order);
    }

    void draw(Rollout &r, std::size_t index, int col, const double *injected) {
        double *dst = r.noise.data() + (std::size_t)col * m_control_dof;
        if (injected) {
            const double *src = injected + (index * m_step_count + col) * m_control_dof;
            for (int d = 0; d < m_control_dof; d++) dst[d] = src[d];
        } else {
            m_gaussian.sample(dst);
        }
    }

    // mppi.cpp:189-270
    void sample(double time, const double *injected) {
        const int nu = m_control_dof, T = m_step_count;
        m_shift_by = (std::int64_t)((time - m_last_shift_time) / m_time_step);
        if (m_shift_by > 0) {
            m_last_shift_time = time;
            m_shifted = T - m_shift_by;
            // leftCols(shifted) = rightCols(shifted); rightCols(shift_by) = last column replicated
            for (std::int64_t c = 0; c < m_shifted; c++)
                for (int d = 0; d < nu; d++) m_optimal_control_shifted[c * nu + d] = m_optimal_control[(c + m_shift_by) * nu + d];
            for (std::int64_t c = m_shifted; c < T; c++)
                for (int d = 0; d < nu; d++) m_optimal_control_shifted[c * nu + d] = m_optimal_control[(std::size_t)(T - 1) * nu + d];
        }
        std::iota(m_ordered_rollouts.begin(), m_ordered_rollouts.end(), (std::size_t)s_static_rollouts);
        // Sort keys are the PREVIOUS update's costs. NaN is ordered as +inf (the reference's
        // comparator is not a strict weak order for NaN — undefined there; SURVEY A-2).
        auto key = [this](std::size_t i) { double c = m_rollouts[i].cost; return std::isnan(c) ? std::numeric_limits<double>::infinity() : c; };
        std::stable_sort(m_ordered_rollouts.begin(), m_ordered_rollouts.end(), [&](std::size_t l, std::size_t r) { return key(l) < key(r); });
        std::size_t keep = std::min(m_keep_best_rollouts, m_ordered_rollouts.size());
        if (m_shift_by > 0) {
            for (std::size_t n = 0; n < keep; n++) {
                std::size_t index = m_ordered_rollouts[n];
                Rollout &r = m_rollouts[index];
                for (std::int64_t c = 0; c < m_shifted; c++)
                    for (int d = 0; d < nu; d++) r.noise[c * nu + d] = r.noise[(c + m_shift_by) * nu + d];
                for (int c = (int)m_shifted; c < T; c++) draw(r, index, c, injected);
            }
        }
        for (std::size_t n = keep; n < m_ordered_rollouts.size(); n++) {
            std::size_t index = m_ordered_rollouts[n];
            for (int c = 0; c < T; c++) draw(m_rollouts[index], index, c, injected);
        }
        // rollout 0 stays zero; rollout 1 = -(unshifted previous optimum) (mppi.cpp:269)
        for (std::size_t i = 0; i < m_optimal_control.size(); i++) m_rollouts[1].noise[i] = -m_optimal_control[i];
    }

    // mppi.cpp:272-307: static contiguous partition over `threads` workers, join as barrier.
    void rollout() {
        int each = m_rollout_count / (int)m_thread_count, distribute = m_rollout_count % (int)m_thread_count;
        int start = 0;
        for (unsigned th = 0; th < m_thread_count; th++) {
            int stop = start + each;
            if (distribute > 0) { stop += 1; distribute -= 1; }
            if (start == stop) break;
            auto body = [this, th, start, stop]() { for (int i = start; i < stop; i++) rollout_one(&m_rollouts[i], m_dynamics[th].get(), m_cost[th].get()); };
            if (m_thread_count == 1) body(); else m_pool->submit(th, body);
            start = stop;
        }
        if (m_thread_count > 1) { m_pool->launch(); m_pool->wait(); }
    }

    // mppi.cpp:309-342
    void rollout_one(Rollout *r, Dynamics *dynamics, Cost *cost) {
        const int nu = m_control_dof;
        std::vector<double> state = m_rollout_state, control(nu);
        dynamics->set_state(state.data(), m_rollout_time);
        cost->reset(m_rollout_time);
        r->cost = 0.0;
        for (int step = 0; step < m_step_count; ++step) {
            for (int d = 0; d < nu; d++) control[d] = m_optimal_control_shifted[(std::size_t)step * nu + d] + r->noise[(std::size_t)step * nu + d];
            double step_cost = std::pow(m_cost_discount_factor, step) * cost->get_cost(state.data(), control.data(), dynamics, m_rollout_time + step * m_time_step);
            if (std::isnan(step_cost)) { r->cost = NAN; return; }
            r->cost += step_cost;
            const double *next = dynamics->step(control.data(), m_time_step);
            state.assign(next, next + m_state_dof);
        }
    }

    // mppi.cpp:344-448
    void optimise() {
        const int nu = m_control_dof, T = m_step_count;
        // minmax_element over the non-NaN costs: first smallest, last largest.
        bool any = false; std::size_t imin = 0, imax = 0, valid = 0;
        for (std::size_t i = 0; i < (std::size_t)m_rollout_count; i++) {
            double c = m_rollouts[i].cost;
            if (std::isnan(c)) continue;
            valid++;
            if (!any) { any = true; imin = imax = i; continue; }
            if (c < m_rollouts[imin].cost) imin = i;
            if (!(c < m_rollouts[imax].cost)) imax = i;
        }
        if (!any || imin == imax) throw std::runtime_error("all nan rollouts");
        (void)valid;
        double minimum = m_rollouts[imin].cost, maximum = m_rollouts[imax].cost;
        m_argmin = imin; m_min = minimum; m_max = maximum;
        double difference = maximum - minimum;
        if (difference < 1e-6) return;
        double total = 0.0;
        for (std::size_t i = 0; i < (std::size_t)m_rollout_count; ++i) {
            double cost = m_rollouts[i].cost;
            if (std::isnan(cost)) { m_weights[i] = 0.0; continue; }
            double likelihood = std::exp(-m_cost_scale * (cost - minimum) / difference);
            total += likelihood;
            m_weights[i] = likelihood;
        }
        for (auto &w : m_weights) w = w / total;
        for (std::size_t e = 0; e < m_gradient.size(); e++) m_gradient[e] = m_rollouts[0].noise[e] * m_weights[0];
        for (std::size_t i = 1; i < (std::size_t)m_rollout_count; ++i) {
            const double w = m_weights[i]; const double *n = m_rollouts[i].noise.data();
            for (std::size_t e = 0; e < m_gradient.size(); e++) m_gradient[e] += n[e] * w;
        }
        for (std::size_t e = 0; e < m_gradient.size(); e++) m_optimal_control_shifted[e] += m_gradient[e] * m_gradient_step;
        if (m_smoothing) {
            m_smoothing->reset(m_rollout_time);
            for (int i = 0; i < T; i++) m_smoothing->add_measurement(&m_optimal_control_shifted[(std::size_t)i * nu], m_rollout_time + i * m_time_step);
            for (int i = 0; i < T; i++) m_smoothing->apply(&m_optimal_control_shifted[(std::size_t)i * nu], m_rollout_time + i * m_time_step);
        }
        if (m_bound_control) {
            for (int i = 0; i < T; i++)
                for (int d = 0; d < nu; d++) {
                    double &u = m_optimal_control_shifted[(std::size_t)i * nu + d];
                    u = std::max(std::min(u, m_control_max[d]), m_control_min[d]);
                }
        }
    }

    // mppi.cpp:450-479 (m_filter is always nullptr in the product, actor.cpp:100)
    void filter() {
        const int nu = m_control_dof;
        std::vector<double> state = m_rollout_state;
        Dynamics *dynamics = m_dynamics[0].get();
        Cost *cost = m_cost[0].get();
        dynamics->set_state(state.data(), m_rollout_time);
        cost->reset(m_rollout_time);
        m_optimal_cost = 0.0;
        for (int step = 0; step < m_step_count; ++step) {
            const double *control = &m_optimal_control_shifted[(std::size_t)step * nu];
            double step_cost = std::pow(m_cost_discount_factor, step) * cost->get_cost(state.data(), control, dynamics, m_rollout_time + step * m_time_step);
            m_optimal_cost += step_cost;
            const double *next = dynamics->step(control, m_time_step);
            state.assign(next, next + m_state_dof);
        }
    }

public:
    std::size_t m_argmin = 0; double m_min = 0.0, m_max = 0.0;
    const std::vector<std::size_t> &ordered() const { return m_ordered_rollouts; }

private:
    const int m_step_count;
    const double m_time_step;
    const int m_rollout_count;
    const unsigned m_thread_count;
    const int m_state_dof, m_control_dof;
    double m_update_last = 0, m_update_duration = 0;
    std::size_t m_update_count = 0;
    std::unique_ptr<Pool> m_pool;
    std::vector<std::unique_ptr<Dynamics>> m_dynamics;
    std::vector<std::unique_ptr<Cost>> m_cost;
    Gaussian m_gaussian;
    std::vector<double> m_rollout_state;
    double m_rollout_time = 0.0;
    double m_last_rollout_time = 0.0;  // uninitialised in the reference (mppi.hpp:594)
    double m_last_shift_time = 0.0;
    std::int64_t m_shift_by = 0, m_shifted = 0;
    const double m_cost_discount_factor, m_cost_scale;
    std::vector<Rollout> m_rollouts;
    std::vector<double> m_weights, m_gradient;
    const double m_gradient_step;
    std::vector<double> m_optimal_control_shifted, m_optimal_control;
    double m_optimal_cost = 0.0;
    const std::size_t m_keep_best_rollouts;
    std::vector<std::size_t> m_ordered_rollouts;
    std::optional<SgFilter> m_smoothing;
    const bool m_bound_control;
    std::vector<double> m_control_min, m_control_max;
    std::optional<std::vector<double>> m_control_default;
};

}  // namespace oracle
