// TEST INFRASTRUCTURE — drives the REFERENCE's own mppi::Trajectory (src/controller/mppi.cpp,
// filter.cpp, gaussian.hpp, gram_savitzky_golay.cpp compiled unmodified from /root/reference by
// oracle/Makefile `ref`, against oracle/ref_shim) so the restatement in mppi_oracle.hpp can be
// pinned against outputs of the reference itself. Only valid for rollouts + 2 <= 255 (the
// reference's std::uint8_t indices, mppi.hpp:639,642). The robot dynamics and objectives plugged
// in here are the oracle's (pinocchio is absent) — what this library pins is the controller:
// sampling order, warm start, weighting, gradient step, smoothing, clamping, readout.
#include <cstring>
#include <memory>

#include "controller/mppi.hpp"   // the reference's header
#include "logging/mppi.hpp"      // the reference's CSV logger
#include "systems.hpp"           // oracle systems

namespace {

struct RefToyDynamics : mppi::Dynamics {
    VectorXd x = VectorXd(4);
    std::unique_ptr<mppi::Dynamics> copy() override { return std::make_unique<RefToyDynamics>(*this); }
    Eigen::Ref<VectorXd> step(const VectorXd &u, double dt) override {
        x[2] += u[0] * dt; x[3] += u[1] * dt; x[0] += x[2] * dt; x[1] += x[3] * dt;
        return x;
    }
    void set_state(const VectorXd &s, double) override { x = s; }
    Eigen::Ref<VectorXd> get_state() override { return x; }
    constexpr int get_control_dof() override { return 2; }
    constexpr int get_state_dof() override { return 4; }
};

struct RefToyCost : mppi::Cost {
    oracle::ToyCost inner;
    explicit RefToyCost(const mppi_b200_toy_objective &p) : inner(p) {}
    std::unique_ptr<mppi::Cost> copy() override { return std::make_unique<RefToyCost>(*this); }
    void reset(double) override {}
    double get_cost(const VectorXd &s, const VectorXd &u, mppi::Dynamics *, double t) override { return inner.get_cost(s.data(), u.data(), nullptr, t); }
    constexpr int get_control_dof() override { return 2; }
    constexpr int get_state_dof() override { return 4; }
};

struct RefFrankaDynamics : mppi::Dynamics {
    oracle::FrankaDynamics inner;
    VectorXd x = VectorXd(31);
    std::unique_ptr<mppi::Dynamics> copy() override { return std::make_unique<RefFrankaDynamics>(*this); }
    Eigen::Ref<VectorXd> step(const VectorXd &u, double dt) override {
        const double *n = inner.step(u.data(), dt);
        std::copy(n, n + 31, x.data());
        return x;
    }
    void set_state(const VectorXd &s, double t) override { inner.set_state(s.data(), t); x = s; }
    Eigen::Ref<VectorXd> get_state() override { return x; }
    constexpr int get_control_dof() override { return 12; }
    constexpr int get_state_dof() override { return 31; }
};

template <class Inner> struct RefFrankaCost : mppi::Cost {
    Inner inner;
    explicit RefFrankaCost(Inner i) : inner(std::move(i)) {}
    std::unique_ptr<mppi::Cost> copy() override { return std::make_unique<RefFrankaCost>(*this); }
    void reset(double t) override { inner.reset(t); }
    double get_cost(const VectorXd &s, const VectorXd &u, mppi::Dynamics *d, double t) override {
        return inner.get_cost(s.data(), u.data(), &static_cast<RefFrankaDynamics *>(d)->inner, t);
    }
    constexpr int get_control_dof() override { return 12; }
    constexpr int get_state_dof() override { return 31; }
};

struct Handle {
    std::unique_ptr<mppi::Trajectory> traj;
    std::shared_ptr<oracle::WrenchTable> table;
    double dt;
};

}  // namespace

extern "C" {

void *ref_create(const mppi_b200_config *c, const void *params, size_t) {
    if (c->rollouts + 2 > 255) return nullptr;  // std::uint8_t loop counters never terminate (mppi.cpp:381)
    mppi::Configuration cfg;
    cfg.initial_state = VectorXd(c->state_dof);
    cfg.rollouts = c->rollouts;
    cfg.keep_best_rollouts = c->keep_best_rollouts;
    cfg.time_step = c->time_step; cfg.horison = c->horison; cfg.gradient_step = c->gradient_step;
    cfg.cost_scale = c->cost_scale; cfg.cost_discount_factor = c->cost_discount_factor;
    cfg.covariance = MatrixXd(c->covariance_rows, c->covariance_cols);
    std::copy(c->covariance, c->covariance + (size_t)c->covariance_rows * c->covariance_cols, cfg.covariance.data());
    cfg.control_bound = c->control_bound != 0;
    cfg.control_min = VectorXd(c->control_limits_size); cfg.control_max = VectorXd(c->control_limits_size);
    std::copy(c->control_min, c->control_min + c->control_limits_size, cfg.control_min.data());
    std::copy(c->control_max, c->control_max + c->control_limits_size, cfg.control_max.data());
    if (c->control_default) { VectorXd d(c->control_dof); std::copy(c->control_default, c->control_default + c->control_dof, d.data()); cfg.control_default = d; }
    if (c->smoothing) cfg.smoothing = mppi::Configuration::Smoothing{c->smoothing_window, c->smoothing_order};
    cfg.threads = (unsigned)c->threads;

    auto h = std::make_unique<Handle>();
    h->dt = c->time_step;
    std::unique_ptr<mppi::Dynamics> dyn; std::unique_ptr<mppi::Cost> cost;
    if (c->system == MPPI_B200_SYSTEM_TOY) {
        dyn = std::make_unique<RefToyDynamics>();
        cost = std::make_unique<RefToyCost>(*static_cast<const mppi_b200_toy_objective *>(params));
    } else if (c->objective == MPPI_B200_OBJECTIVE_TRACK_POINT) {
        dyn = std::make_unique<RefFrankaDynamics>();
        cost = std::make_unique<RefFrankaCost<oracle::TrackPointCost>>(oracle::TrackPointCost(*static_cast<const mppi_b200_track_point *>(params)));
    } else {
        dyn = std::make_unique<RefFrankaDynamics>();
        h->table = std::make_shared<oracle::WrenchTable>();
        cost = std::make_unique<RefFrankaCost<oracle::AssistedManipulationCost>>(
            oracle::AssistedManipulationCost(*static_cast<const mppi_b200_assisted_manipulation *>(params), h->table));
    }
    h->traj = mppi::Trajectory::create(cfg, std::move(dyn), std::move(cost));
    if (!h->traj) return nullptr;
    return h.release();
}

void ref_destroy(void *p) { delete static_cast<Handle *>(p); }

int ref_update(void *p, const double *state, double time, const double *wrench) {
    auto *h = static_cast<Handle *>(p);
    VectorXd s(h->traj->get_state_dof());
    std::copy(state, state + s.size(), s.data());
    if (h->table) {
        h->table->present = wrench != nullptr; h->table->t0 = time; h->table->dt = h->dt;
        if (wrench) h->table->w.assign(wrench, wrench + (size_t)h->traj->get_step_count() * 6);
    }
    try { h->traj->update(s, time); } catch (const std::exception &) { return -4; }
    return 0;
}

int ref_get(void *p, double *control, double time) {
    auto *h = static_cast<Handle *>(p);
    VectorXd c(h->traj->get_control_dof());
    h->traj->get(c, time);
    std::copy(c.data(), c.data() + c.size(), control);
    return 0;
}

int ref_read(void *p, int what, double *out, size_t bytes) {
    auto &t = *static_cast<Handle *>(p)->traj;
    size_t K = t.get_rollout_count(), n = (size_t)t.get_control_dof() * t.get_step_count();
    switch (what) {
        case MPPI_B200_READ_OPTIMAL: if (bytes != n * 8) return -1; std::memcpy(out, t.trajectory().data(), bytes); return 0;
        case MPPI_B200_READ_COSTS: if (bytes != K * 8) return -1; for (size_t k = 0; k < K; k++) out[k] = t.get_rollouts()[k].cost; return 0;
        case MPPI_B200_READ_WEIGHTS: if (bytes != K * 8) return -1; std::memcpy(out, t.get_weights().data(), bytes); return 0;
        case MPPI_B200_READ_GRADIENT: if (bytes != n * 8) return -1; std::memcpy(out, t.get_gradient().data(), bytes); return 0;
        case MPPI_B200_READ_NOISE: if (bytes != K * n * 8) return -1; for (size_t k = 0; k < K; k++) std::memcpy(out + k * n, t.get_rollouts()[k].noise.data(), n * 8); return 0;
        case MPPI_B200_READ_OPTIMAL_COST: if (bytes != 8) return -1; out[0] = t.get_optimal_total_cost(); return 0;
    }
    return -1;
}

// gram_sg::ComputeWeights straight from the reference (gram_savitzky_golay.cpp:46-53)
void ref_sg_weights(int m, int t, int n, int s, double *out) { auto w = gram_sg::ComputeWeights(m, t, n, s); std::memcpy(out, w.data(), w.size() * 8); }

// The reference's SavitzkyGolayFilter driven like mppi.cpp:424-440 on one channel.
void ref_sg_run(int steps, int window, unsigned order, int updates, const double *t0s, double dt, const double *u, double *out) {
    SavitzkyGolayFilter f(steps, 1, window, order, 0, dt);
    for (int n = 0; n < updates; n++) {
        f.reset(t0s[n]);
        MatrixXd m(1, steps);
        for (int i = 0; i < steps; i++) m(0, i) = u[(size_t)n * steps + i];
        for (int i = 0; i < steps; i++) f.add_measurement(m.col(i), t0s[n] + i * dt);
        for (int i = 0; i < steps; i++) f.apply(m.col(i), t0s[n] + i * dt);
        for (int i = 0; i < steps; i++) out[(size_t)n * steps + i] = m(0, i);
    }
}

// ---- the reference's own CSV logger over its own Trajectory (logging/mppi.cpp, csv.hpp, file.hpp), SURVEY §8f-4 ----
void *ref_logger_create(const char *folder, unsigned control_dof, size_t rollouts) {
    logger::MPPI::Configuration c;
    c.folder = folder; c.state_dof = 0; c.control_dof = control_dof; c.rollouts = rollouts;
    return logger::MPPI::create(c).release();
}
void ref_logger_log(void *lg, void *p) { static_cast<logger::MPPI *>(lg)->log(*static_cast<Handle *>(p)->traj); }
void ref_logger_destroy(void *lg) { delete static_cast<logger::MPPI *>(lg); }

}  // extern "C"
