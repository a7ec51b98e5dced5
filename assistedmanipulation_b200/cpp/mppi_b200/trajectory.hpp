// mppi::Trajectory on a B200 — same public API as the reference controller
// (src/controller/mppi.hpp:30-249 plug-in interfaces and Configuration, :267-474 Trajectory),
// implemented over the C ABI of include/mppi_b200.h. Header-only; link with libmppi_b200.so.
//
// What differs from the reference, by necessity:
//   * Cost::get_cost reaches back into the Dynamics object through virtual calls
//     (assisted_manipulation.cpp:46), which cannot cross to a GPU. create() therefore recognises the
//     concrete (Dynamics, Cost) pair — the classes of mppi_b200/systems.hpp implement DeviceBound —
//     and returns nullptr with a reason on std::cerr for anything else. There is no CPU fallback.
//   * get_rollouts() / get_weights() / get_gradient() read back from the device on demand.
//   * the optimal re-rollout (Trajectory::filter, mppi.cpp:450-479) runs on a side stream;
//     get_optimal_total_cost() waits for it.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <algorithm>
#include <limits>
#include <numeric>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "mppi_b200.h"
#include "mppi_b200/linalg.hpp"

namespace mppi {

template <class T> using Ref = mppi_b200::Ref<T>;

// mppi.hpp:30-85
class Dynamics {
public:
    virtual ~Dynamics() = default;
    virtual std::unique_ptr<Dynamics> copy() = 0;
    virtual Ref<VectorXd> step(const VectorXd &control, double dt) = 0;
    virtual void set_state(const VectorXd &state, double time) = 0;
    virtual Ref<VectorXd> get_state() = 0;
    virtual int get_control_dof() = 0;
    virtual int get_state_dof() = 0;
};

// mppi.hpp:93-145
class Cost {
public:
    virtual ~Cost() = default;
    virtual std::unique_ptr<Cost> copy() = 0;
    virtual void reset(double time) = 0;
    virtual double get_cost(const VectorXd &state, const VectorXd &control, Dynamics *dynamics, double time) = 0;
    virtual int get_control_dof() = 0;
    virtual int get_state_dof() = 0;
};

// mppi.hpp:150-176. The product passes nullptr (actor.cpp:100); a non-null filter is refused because it
// would have to run per step inside the optimal re-rollout on the device.
class Filter {
public:
    virtual ~Filter() = default;
    virtual VectorXd filter(Ref<VectorXd> state, Ref<VectorXd> control, double time) = 0;
    virtual void reset(Ref<VectorXd> state, double time) = 0;
};

// mppi.hpp:181-249 (field names verbatim, including `horison`)
struct Configuration {
    VectorXd initial_state;
    std::int64_t rollouts = 1;
    std::int64_t keep_best_rollouts = 0;
    double time_step = 0.01;
    double horison = 1.0;
    double gradient_step = 1.0;
    double cost_scale = 1.0;
    double cost_discount_factor = 1.0;
    MatrixXd covariance;
    bool control_bound = false;
    VectorXd control_min;
    VectorXd control_max;
    std::optional<VectorXd> control_default;
    struct Smoothing { unsigned int window; unsigned int order; };
    std::optional<Smoothing> smoothing;
    unsigned int threads = 1;
    // device binding (not in the reference): arithmetic and evaluation mode of the rollout kernel
    int precision = MPPI_B200_FP64;
    int dynamics_mode = MPPI_B200_DYNAMICS_FUSED;
    int device = 0;
};

}  // namespace mppi

namespace mppi_b200 {

// Implemented by the Dynamics / Cost classes a device kernel exists for.
struct DeviceBoundDynamics {
    virtual ~DeviceBoundDynamics() = default;
    virtual int device_system() const = 0;
    // forecast->get_end_effector_wrench(time).head(6) (dynamics.hpp:275-278); false = no forecast handle
    virtual bool forecast_wrench(double /*time*/, double * /*wrench6*/) const { return false; }
    // the same T x 6 table produced on the device by the forecast producer (forecast.hpp), or nullptr
    virtual const double *forecast_table_device(double /*time*/, double /*time_step*/, int /*steps*/) const { return nullptr; }
};
struct DeviceBoundCost {
    virtual ~DeviceBoundCost() = default;
    virtual int device_objective() const = 0;
    virtual const void *device_params(std::size_t *size) const = 0;
    // per-term totals of the optimal re-rollout, as logging/assisted_manipulation.cpp:58-103 reads them
    virtual void set_optimal_breakdown(const double * /*terms8*/) {}
};

}  // namespace mppi_b200

namespace mppi_b200 {

// controller/gaussian.hpp:12-90 restated for the reference-RNG mode of the facade: std::mt19937 with its
// default seed, std::normal_distribution<double>, transform = V * sqrt(Lambda) of the covariance with the
// eigenvalues in ascending order (SelfAdjointEigenSolver). Host code: it only feeds the injected-noise
// path, so that closed-loop traces can be compared with a real reference build sample for sample.
class ReferenceGaussian {
public:
    ReferenceGaussian(int n, const double *covariance /* n x n column-major */) : m_n(n), m_transform((std::size_t)n * n, 0.0) {
        std::vector<double> a(covariance, covariance + (std::size_t)n * n), V((std::size_t)n * n, 0.0), ev((std::size_t)n);
        auto A = [&](int r, int c) -> double & { return a[(std::size_t)c * n + r]; };
        auto E = [&](int r, int c) -> double & { return V[(std::size_t)c * n + r]; };
        for (int i = 0; i < n; i++) E(i, i) = 1.0;
        for (int sweep = 0; sweep < 64; sweep++) {  // cyclic Jacobi; exact for the diagonal covariances of base.hpp:79-83
            double off = 0.0;
            for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A(p, q) * A(p, q);
            if (off == 0.0) break;
            for (int p = 0; p < n; p++)
                for (int q = p + 1; q < n; q++) {
                    if (A(p, q) == 0.0) continue;
                    const double theta = (A(q, q) - A(p, p)) / (2.0 * A(p, q));
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                    const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                    for (int k = 0; k < n; k++) { const double x = A(k, p), y = A(k, q); A(k, p) = c * x - s * y; A(k, q) = s * x + c * y; }
                    for (int k = 0; k < n; k++) { const double x = A(p, k), y = A(q, k); A(p, k) = c * x - s * y; A(q, k) = s * x + c * y; }
                    for (int k = 0; k < n; k++) { const double x = E(k, p), y = E(k, q); E(k, p) = c * x - s * y; E(k, q) = s * x + c * y; }
                }
        }
        for (int i = 0; i < n; i++) ev[(std::size_t)i] = A(i, i);
        for (int i = 0; i < n - 1; i++) {  // Eigen sorts with a selection sort that swaps
            int k = 0;
            for (int j = 1; j < n - i; j++) if (ev[(std::size_t)(i + j)] < ev[(std::size_t)(i + k)]) k = j;
            if (k > 0) { std::swap(ev[(std::size_t)i], ev[(std::size_t)(i + k)]); for (int r = 0; r < n; r++) std::swap(E(r, i), E(r, i + k)); }
        }
        for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) m_transform[(std::size_t)c * n + r] = E(r, c) * std::sqrt(ev[(std::size_t)c]);
    }
    // gaussian.hpp:70-75: mean + transform * z, z drawn element by element
    void operator()(double *out) {
        std::vector<double> z((std::size_t)m_n);
        for (auto &v : z) v = m_distribution(m_generator);
        for (int r = 0; r < m_n; r++) out[r] = 0.0;
        for (int c = 0; c < m_n; c++) for (int r = 0; r < m_n; r++) out[r] += m_transform[(std::size_t)c * m_n + r] * z[(std::size_t)c];
    }
private:
    int m_n;
    std::vector<double> m_transform;
    std::mt19937 m_generator;
    std::normal_distribution<double> m_distribution{0, 1};
};

}  // namespace mppi_b200

namespace mppi {

class Trajectory {
public:
    // mppi.hpp:275-302
    class Rollout {
    public:
        MatrixXd noise;
        double cost;
    private:
        friend class Trajectory;
        Rollout(std::size_t control_dof, std::size_t steps) : noise((std::ptrdiff_t)control_dof, (std::ptrdiff_t)steps), cost(0.0) {}
    };

    static const constexpr std::int64_t s_static_rollouts = 2;

    // mppi.cpp:11-77: nullptr + reason on std::cerr on failure
    static std::unique_ptr<Trajectory> create(const Configuration &configuration, std::unique_ptr<Dynamics> &&dynamics, std::unique_ptr<Cost> &&cost,
                                              std::unique_ptr<Filter> &&filter = nullptr) {
        auto *dd = dynamic_cast<mppi_b200::DeviceBoundDynamics *>(dynamics.get());
        auto *dc = dynamic_cast<mppi_b200::DeviceBoundCost *>(cost.get());
        if (!dd || !dc) { std::cerr << "mppi_b200: no device implementation for this (dynamics, cost) pair; there is no CPU fallback" << std::endl; return nullptr; }
        if (filter) { std::cerr << "mppi_b200: a per-step mppi::Filter is not supported on the device" << std::endl; return nullptr; }
        mppi_b200_config c{};
        c.abi_version = MPPI_B200_ABI_VERSION;
        c.system = dd->device_system(); c.objective = dc->device_objective();
        c.precision = configuration.precision; c.dynamics_mode = configuration.dynamics_mode; c.device = configuration.device;
        c.rank = 0; c.world_size = 1;
        // the C ABI repeats the reference's checks (mppi.cpp:18-69) with the same messages
        c.state_dof = dynamics->get_state_dof() == cost->get_state_dof() ? dynamics->get_state_dof() : -cost->get_state_dof();
        c.control_dof = dynamics->get_control_dof();
        if (dynamics->get_control_dof() != cost->get_control_dof()) {
            std::cerr << "controller dynamics control dof " << dynamics->get_control_dof() << " != cost control dof " << cost->get_control_dof() << std::endl;
            return nullptr;
        }
        if (dynamics->get_state_dof() != cost->get_state_dof()) {
            std::cerr << "controller dynamics state dof " << dynamics->get_state_dof() << " != cost state dof " << cost->get_state_dof() << std::endl;
            return nullptr;
        }
        c.rollouts = configuration.rollouts; c.keep_best_rollouts = configuration.keep_best_rollouts;
        c.time_step = configuration.time_step; c.horison = configuration.horison; c.gradient_step = configuration.gradient_step;
        c.cost_scale = configuration.cost_scale; c.cost_discount_factor = configuration.cost_discount_factor;
        c.covariance = configuration.covariance.data(); c.covariance_rows = (int)configuration.covariance.rows(); c.covariance_cols = (int)configuration.covariance.cols();
        c.control_bound = configuration.control_bound;
        c.control_limits_size = (int)configuration.control_min.size() == (int)configuration.control_max.size() ? (int)configuration.control_min.size() : -1;
        c.control_min = configuration.control_min.data(); c.control_max = configuration.control_max.data();
        c.control_default = configuration.control_default ? configuration.control_default->data() : nullptr;
        c.smoothing = configuration.smoothing ? 1 : 0;
        c.smoothing_window = configuration.smoothing ? configuration.smoothing->window : 0;
        c.smoothing_order = configuration.smoothing ? configuration.smoothing->order : 0;
        c.threads = (int)configuration.threads;
        std::size_t psize = 0;
        const void *params = dc->device_params(&psize);
        mppi_b200_engine *engine = nullptr;
        if (mppi_b200_create(&c, params, psize, &engine) != MPPI_B200_OK) { std::cerr << mppi_b200_last_error(nullptr) << std::endl; return nullptr; }
        return std::unique_ptr<Trajectory>(new Trajectory(configuration, engine, std::move(dynamics), std::move(cost)));
    }

    ~Trajectory() { mppi_b200_destroy(m_engine); }
    Trajectory(const Trajectory &) = delete;
    Trajectory &operator=(const Trajectory &) = delete;

    // mppi.cpp:154-187
    void update(const Ref<VectorXd> state, double time) {
        for (int i = 0; i < m_state_dof; i++) m_rollout_state[i] = state[i];
        const double *wrench = nullptr;
        const double *device_table = m_device_dynamics->forecast_table_device(time, m_time_step, m_step_count);
        if (mppi_b200_set_wrench_device(m_engine, device_table) != MPPI_B200_OK && device_table) throw std::runtime_error(mppi_b200_last_error(m_engine));
        bool have = device_table == nullptr;
        for (int k = 0; k < m_step_count && have; k++) have = m_device_dynamics->forecast_wrench(time + k * m_time_step, &m_wrench[(std::size_t)6 * k]);
        if (have && m_step_count > 0) wrench = m_wrench.data();
        if (m_reference_rng) sample_like_the_reference(time);
        const void *noise = m_reference_rng ? m_reference_noise.data() : m_injected;
        const int source = noise ? MPPI_B200_NOISE_HOST : MPPI_B200_NOISE_PHILOX;
        // get() may run concurrently (mppi.cpp:179,492): the engine takes its own lock around the publication of the new
        // sequence only, like the reference around `m_optimal_control = m_optimal_control_shifted` (mppi.cpp:178-182) —
        // a control thread calling get() never waits for a running update
        const int rc = mppi_b200_update(m_engine, m_rollout_state.data(), time, wrench, noise, source, m_seed);
        if (rc == MPPI_B200_ERR_ALL_NAN) throw std::runtime_error("all nan rollouts");          // mppi.cpp:370
        if (rc == MPPI_B200_ERR_TIME) throw std::runtime_error(mppi_b200_last_error(m_engine));   // filter.cpp:37-44
        if (rc != MPPI_B200_OK) throw std::runtime_error(std::string("mppi_b200: ") + mppi_b200_last_error(m_engine));
        mppi_b200_last_update_device_seconds(m_engine, &m_update_duration);
        m_update_last = time;
        ++m_update_count;
        m_stale = true;
    }

    inline unsigned int get_state_dof() const { return m_state_dof; }
    inline unsigned int get_control_dof() const { return m_control_dof; }
    inline double get_time_step() const { return m_time_step; }
    inline unsigned int get_step_count() const { return m_step_count; }
    inline double get_update_duration() const { return m_update_duration; }
    inline double get_update_last() const { return m_update_last; }
    inline std::size_t get_update_count() const { return m_update_count; }
    inline std::size_t get_rollout_count() const { return m_rollout_count; }
    inline const auto &get_rolled_out_state() const { return m_rollout_state; }

    // const like the reference's (mppi.hpp:408-431) so loggers taking `const Trajectory &` work; the host copies are caches
    inline const VectorXd &get_weights() const { refresh(); return m_weights; }
    inline const MatrixXd &get_gradient() const { refresh(); return m_gradient; }
    inline const std::vector<Rollout> &get_rollouts() const { refresh_rollouts(); return m_rollouts; }
    inline const MatrixXd &get_optimal_rollout() const { refresh(); return m_optimal_control; }
    inline const MatrixXd &trajectory() const { refresh(); return m_optimal_control; }
    inline double get_optimal_total_cost() const { double c = 0; mppi_b200_read(m_engine, MPPI_B200_READ_OPTIMAL_COST, &c, sizeof c); return c; }
    inline const Cost &get_optimal_cost() const {
        double bd[8];
        if (mppi_b200_read(m_engine, MPPI_B200_READ_BREAKDOWN, bd, sizeof bd) == MPPI_B200_OK) m_device_cost->set_optimal_breakdown(bd);
        return *m_cost;
    }
    inline const Dynamics &get_optimal_dynamics() const { return *m_dynamics; }

    // mppi.cpp:481-512
    void get(Ref<VectorXd> control, double time) {
        if (mppi_b200_get(m_engine, control.data(), time) != MPPI_B200_OK) throw std::logic_error("time >= m_last_rollout_time");
    }
    inline VectorXd operator()(double time) { VectorXd control(m_control_dof); get(control, time); return control; }

    // ---- device binding extras (not in the reference) ----------------------------------------------
    // Injected-noise mode for equivalence testing: [(K+2)][T][nu] doubles, used wherever the reference
    // would draw a fresh column; nullptr returns to in-kernel Philox.
    void set_injected_noise(const double *noise) { m_injected = noise; }
    void set_seed(std::uint64_t seed) { m_seed = seed; }
    // Reference-RNG mode: the noise is drawn on the host with the reference's generator in the reference's
    // order (kept rollouts' new tail columns first, then every resampled rollout, both in cost order,
    // mppi.cpp:222-262) and injected; the device then reproduces a reference build sample for sample.
    void use_reference_rng(bool on) { m_reference_rng = on; }
    mppi_b200_engine *engine() { return m_engine; }

private:
    Trajectory(const Configuration &configuration, mppi_b200_engine *engine, std::unique_ptr<Dynamics> &&dynamics, std::unique_ptr<Cost> &&cost)
        : m_engine(engine), m_dynamics(std::move(dynamics)), m_cost(std::move(cost)),
          m_device_dynamics(dynamic_cast<mppi_b200::DeviceBoundDynamics *>(m_dynamics.get())),
          m_device_cost(dynamic_cast<mppi_b200::DeviceBoundCost *>(m_cost.get())),
          m_step_count((int)std::ceil(configuration.horison / configuration.time_step)), m_time_step(configuration.time_step),
          m_rollout_count((int)(configuration.rollouts + s_static_rollouts)), m_state_dof(m_dynamics->get_state_dof()), m_control_dof(m_dynamics->get_control_dof()),
          m_rollout_state(m_dynamics->get_state_dof()), m_weights(m_rollout_count), m_gradient(m_control_dof, m_step_count), m_optimal_control(m_control_dof, m_step_count),
          m_wrench((std::size_t)6 * m_step_count, 0.0) {
        m_covariance = configuration.covariance;
        m_keep_best = configuration.keep_best_rollouts;
        m_rollout_state.setZero(); m_weights.setZero(); m_gradient.setZero(); m_optimal_control.setZero();
    }

    void sample_like_the_reference(double time) {
        const std::size_t nu = (std::size_t)m_control_dof, T = (std::size_t)m_step_count, R = (std::size_t)m_rollout_count;
        if (!m_gaussian) m_gaussian = std::make_unique<mppi_b200::ReferenceGaussian>(m_control_dof, m_covariance.data());
        m_reference_noise.assign(R * T * nu, 0.0);
        // mppi.cpp:194-201
        const std::int64_t shift_by = (std::int64_t)((time - m_last_shift_time) / m_time_step);
        if (shift_by > 0) m_last_shift_time = time;
        const std::int64_t shifted = (std::int64_t)T - shift_by;
        // mppi.cpp:222-231: indices 2.. sorted by the PREVIOUS update's costs, stable
        std::vector<double> costs(R, 0.0);
        if (m_update_count > 0) mppi_b200_read(m_engine, MPPI_B200_READ_COSTS, costs.data(), R * sizeof(double));
        std::vector<std::size_t> order(R - 2);
        std::iota(order.begin(), order.end(), (std::size_t)2);
        auto key = [&](std::size_t i) { return std::isnan(costs[i]) ? std::numeric_limits<double>::infinity() : costs[i]; };
        std::stable_sort(order.begin(), order.end(), [&](std::size_t l, std::size_t r) { return key(l) < key(r); });
        const std::size_t keep = std::min<std::size_t>((std::size_t)m_keep_best, order.size());
        if (shift_by > 0)
            for (std::size_t i = 0; i < keep; i++)
                for (std::size_t c = (std::size_t)std::max<std::int64_t>(shifted, 0); c < T; c++) (*m_gaussian)(&m_reference_noise[(order[i] * T + c) * nu]);
        for (std::size_t i = keep; i < order.size(); i++)
            for (std::size_t c = 0; c < T; c++) (*m_gaussian)(&m_reference_noise[(order[i] * T + c) * nu]);
    }

    void refresh() const {
        if (!m_stale) return;
        const std::size_t n = (std::size_t)m_control_dof * m_step_count;
        mppi_b200_read(m_engine, MPPI_B200_READ_OPTIMAL, m_optimal_control.data(), n * sizeof(double));
        mppi_b200_read(m_engine, MPPI_B200_READ_GRADIENT, m_gradient.data(), n * sizeof(double));
        mppi_b200_read(m_engine, MPPI_B200_READ_WEIGHTS, m_weights.data(), (std::size_t)m_rollout_count * sizeof(double));
        m_stale = false;
    }
    void refresh_rollouts() const {
        const std::size_t n = (std::size_t)m_control_dof * m_step_count;
        if (m_rollouts.empty()) m_rollouts.assign((std::size_t)m_rollout_count, Rollout(m_control_dof, m_step_count));
        std::vector<double> noise((std::size_t)m_rollout_count * n), costs((std::size_t)m_rollout_count);
        mppi_b200_read(m_engine, MPPI_B200_READ_NOISE, noise.data(), noise.size() * sizeof(double));
        mppi_b200_read(m_engine, MPPI_B200_READ_COSTS, costs.data(), costs.size() * sizeof(double));
        for (std::size_t k = 0; k < m_rollouts.size(); k++) {
            for (std::size_t e = 0; e < n; e++) m_rollouts[k].noise.data()[e] = noise[k * n + e];
            m_rollouts[k].cost = costs[k];
        }
    }

    mppi_b200_engine *m_engine;
    std::unique_ptr<Dynamics> m_dynamics;
    std::unique_ptr<Cost> m_cost;
    mppi_b200::DeviceBoundDynamics *m_device_dynamics;
    mppi_b200::DeviceBoundCost *m_device_cost;
    const int m_step_count;
    const double m_time_step;
    const int m_rollout_count;
    const int m_state_dof, m_control_dof;
    double m_update_last = 0.0, m_update_duration = 0.0;
    std::size_t m_update_count = 0;
    VectorXd m_rollout_state;
    mutable VectorXd m_weights;
    mutable MatrixXd m_gradient, m_optimal_control;
    mutable std::vector<Rollout> m_rollouts;
    std::vector<double> m_wrench;
    const double *m_injected = nullptr;
    bool m_reference_rng = false;
    std::unique_ptr<mppi_b200::ReferenceGaussian> m_gaussian;
    std::vector<double> m_reference_noise;
    MatrixXd m_covariance;
    std::int64_t m_keep_best = 0;
    double m_last_shift_time = 0.0;
    std::uint64_t m_seed = 0x5EED0000ull;
    mutable bool m_stale = true;
};

}  // namespace mppi
