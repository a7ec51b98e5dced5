// TEST INFRASTRUCTURE — drives the REFERENCE's own mppi::Trajectory (src/controller/mppi.cpp,
// filter.cpp, gaussian.hpp, gram_savitzky_golay.cpp compiled unmodified from /root/reference by
// oracle/Makefile `ref`, against oracle/ref_shim) so the restatement in mppi_oracle.hpp can be
// pinned against outputs of the reference itself. Only valid for rollouts + 2 <= 255 (the
// reference's std::uint8_t indices, mppi.hpp:639,642). The OBJECTIVES are the reference's own too
// (objective/track_point.cpp, objective/assisted_manipulation.cpp, frankaridgeback/dynamics.cpp, state.cpp compiled
// unmodified); only the rigid-body arithmetic behind FrankaRidgeback::Dynamics is the oracle's (pinocchio is absent).
// What this library pins: the controller (sampling order, warm start, weighting, gradient step, smoothing, clamping,
// readout), the cost functors (cost.hpp) and both objectives, term by term.
#include <cstring>
#include <memory>

#include "controller/mppi.hpp"   // the reference's header
#include "logging/mppi.hpp"      // the reference's CSV logger
#include "frankaridgeback/dynamics.hpp"                         // the reference's robot-specific Dynamics interface + DynamicsForecast
#include "frankaridgeback/objective/track_point.hpp"            // the reference's objectives
#include "frankaridgeback/objective/assisted_manipulation.hpp"
#include "controller/energy.hpp"                                // the reference's energy tank
#include "systems.hpp"           // oracle systems
#include "dynamics_forecast_oracle.hpp"   // rotation_to_quaternion, record layout

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

// The reference's robot-specific Dynamics interface (frankaridgeback/dynamics.hpp:416-537) over the oracle's rigid-body
// core — the one layer that cannot come from the reference (pinocchio is absent). Accessors follow
// PinocchioDynamics (pinocchio_dynamics.hpp:112-262): link positions are zero in that backend (:189-192; the
// BODY_COM mode is the RaiSim intent, SURVEY A-3), powers are 0, the simulated wrench is dropped.
struct RefFrankaDynamics : FrankaRidgeback::Dynamics {
    oracle::FrankaDynamics inner;
    VectorXd x = VectorXd(31), jq = VectorXd(12), jv = VectorXd(12);
    FrankaRidgeback::EndEffectorState ee;
    int link_mode = MPPI_B200_LINKS_ZERO;
    std::shared_ptr<oracle::WrenchTable> table;                         // nullptr / !present = no forecast handle
    std::unique_ptr<FrankaRidgeback::DynamicsForecast::Handle> forecast;

    RefFrankaDynamics() = default;
    RefFrankaDynamics(const RefFrankaDynamics &o)
        : inner(o.inner), x(o.x), jq(o.jq), jv(o.jv), ee(o.ee), link_mode(o.link_mode), table(o.table), forecast(o.forecast ? o.forecast->copy() : nullptr) {}
    std::unique_ptr<mppi::Dynamics> copy() override { return std::make_unique<RefFrankaDynamics>(*this); }

    void sync() {
        const auto &c = inner.core;
        for (int i = 0; i < 12; i++) { jq[i] = c.q[i]; jv[i] = c.v[i]; }
        ee.position = Vector3d(c.ee.position.x, c.ee.position.y, c.ee.position.z);
        ee.linear_velocity = Vector3d(c.ee.linear_velocity.x, c.ee.linear_velocity.y, c.ee.linear_velocity.z);
        ee.angular_velocity = Vector3d(c.ee.angular_velocity.x, c.ee.angular_velocity.y, c.ee.angular_velocity.z);
        ee.linear_acceleration = Vector3d(c.ee.linear_acceleration.x, c.ee.linear_acceleration.y, c.ee.linear_acceleration.z);
        ee.angular_acceleration = Vector3d(c.ee.angular_acceleration.x, c.ee.angular_acceleration.y, c.ee.angular_acceleration.z);
        for (int r = 0; r < 6; r++) for (int j = 0; j < 12; j++) ee.jacobian(r, j) = c.ee.jacobian[r][j];
        double xyzw[4];
        oracle::rotation_to_quaternion(c.ee.orientation, xyzw);   // Quaterniond(rotation), pinocchio_dynamics.cpp:219
        ee.orientation = Quaterniond(xyzw[3], xyzw[0], xyzw[1], xyzw[2]);
    }
    Eigen::Ref<VectorXd> step(const VectorXd &u, double dt) override {
        const double *n = inner.step(u.data(), dt);
        std::copy(n, n + 31, x.data());
        sync();
        return x;
    }
    void set_state(const VectorXd &s, double t) override { inner.set_state(s.data(), t); x = s; sync(); }
    Eigen::Ref<VectorXd> get_state() override { return x; }
    constexpr int get_control_dof() override { return 12; }
    constexpr int get_state_dof() override { return 31; }
    const VectorXd &get_joint_position() const override { return jq; }
    const VectorXd &get_joint_velocity() const override { return jv; }
    Vector3d get_frame_position(FrankaRidgeback::Frame frame) override {
        const auto &d = inner.core.data;
        if (frame == FrankaRidgeback::Frame::ARM_MOUNT_JOINT) return Vector3d(d.oMf_mount.p.x, d.oMf_mount.p.y, d.oMf_mount.p.z);
        if (frame == FrankaRidgeback::Frame::PANDA_GRASP_JOINT) return Vector3d(d.oMf_ee.p.x, d.oMf_ee.p.y, d.oMf_ee.p.z);
        std::cerr << "ref_driver: frame " << (int)frame << " is not carried by the oracle core" << std::endl;
        std::abort();
    }
    Quaterniond get_frame_orientation(FrankaRidgeback::Frame) override { return Quaterniond::Identity(); }
    Vector3d get_link_position(FrankaRidgeback::Link link) override {
        if (link_mode == MPPI_B200_LINKS_ZERO) return Vector3d::Zero();
        const auto &p = inner.core.link_com_world[(std::size_t)link];
        return Vector3d(p.x, p.y, p.z);
    }
    const FrankaRidgeback::EndEffectorState &get_end_effector_state() const override { return ee; }
    double get_joint_power() const override { return 0.0; }
    double get_external_power() const override { return 0.0; }
    double get_tank_energy() const override { return inner.core.tank.energy; }
    const FrankaRidgeback::DynamicsForecast::Handle *get_forecast() const override {
        return (forecast && table && table->present) ? forecast.get() : nullptr;
    }
    Vector6d get_end_effector_simulated_wrench() const override { return Vector6d::Zero(); }
    void add_end_effector_simulated_wrench(Vector6d) override {}
};

// What the facade hands the engine: forecast->get_end_effector_wrench(t0 + k dt) tabulated once per update. Served to
// the reference's objective through the reference's own DynamicsForecast (dynamics.hpp:275-278 -> Forecast::forecast).
struct TableForecast : Forecast {
    const oracle::WrenchTable *table;   // raw: the reference's Forecast has no virtual destructor (forecast.hpp:14-56)
    explicit TableForecast(const std::shared_ptr<oracle::WrenchTable> &t) : table(t.get()) {}
    void update(VectorXd, double) override {}
    void update(double) override {}
    VectorXd forecast(double time) override {
        VectorXd w(6);
        const double *row = table->at(time);
        for (int i = 0; i < 6; i++) w[i] = row ? row[i] : 0.0;
        return w;
    }
};
struct TableDynamicsForecast : FrankaRidgeback::DynamicsForecast {   // its constructor is protected (dynamics.hpp:344-349)
    TableDynamicsForecast(const Configuration &c, std::unique_ptr<Forecast> &&f, unsigned steps)
        : DynamicsForecast(c, nullptr, std::move(f), steps) {}
};

FrankaRidgeback::TrackPoint::Configuration track_point_configuration(const mppi_b200_track_point &p) {
    auto c = FrankaRidgeback::TrackPoint::DEFAULT_CONFIGURATION;
    c.point = Vector3d(p.point[0], p.point[1], p.point[2]);
    c.enable_joint_limits = p.enable_joint_limits; c.enable_self_collision_avoidance = p.enable_self_collision_avoidance;
    c.enable_power_limit = p.enable_power_limit; c.enable_reach_limits = p.enable_reach_limits;
    for (int i = 0; i < 12; i++) {
        c.lower_joint_limit[i] = {p.lower_joint_limit[i].bound, p.lower_joint_limit[i].scale, p.lower_joint_limit[i].maximum_cost};
        c.upper_joint_limit[i] = {p.upper_joint_limit[i].bound, p.upper_joint_limit[i].scale, p.upper_joint_limit[i].maximum_cost};
    }
    c.self_collision_limit = {p.self_collision_limit.bound, p.self_collision_limit.scale, p.self_collision_limit.maximum_cost};
    for (int i = 0; i < 8; i++) c.self_collision_radii[i] = p.self_collision_radii[i];
    c.maximum_reach_limit = {p.maximum_reach_limit.bound, p.maximum_reach_limit.scale, p.maximum_reach_limit.maximum_cost};
    return c;
}

FrankaRidgeback::AssistedManipulation::Configuration assisted_configuration(const mppi_b200_assisted_manipulation &p) {
    auto c = FrankaRidgeback::AssistedManipulation::DEFAULT_CONFIGURATION;
    auto left = [](const mppi_b200_barrier &b) { return LeftInverseBarrierFunction{b.bound, b.scale, b.maximum_cost}; };
    auto right = [](const mppi_b200_barrier &b) { return RightInverseBarrierFunction{b.bound, b.scale, b.maximum_cost}; };
    auto quad = [](const mppi_b200_quadratic &q) { return QuadraticCost{q.constant_cost, q.linear_cost, q.quadratic_cost}; };
    c.enable_joint_limit = p.enable_joint_limit; c.enable_self_collision_limit = p.enable_self_collision_limit;
    c.enable_workspace_limit = p.enable_workspace_limit; c.enable_energy_limit = p.enable_energy_limit;
    c.enable_velocity_cost = p.enable_velocity_cost; c.enable_trajectory_cost = p.enable_trajectory_cost;
    c.enable_manipulability_cost = p.enable_manipulability_cost;
    for (int i = 0; i < 12; i++) {
        c.lower_joint_limit[i] = left(p.lower_joint_limit[i]); c.upper_joint_limit[i] = right(p.upper_joint_limit[i]);
        c.velocity_cost[i] = quad(p.velocity_cost[i]);
    }
    c.self_collision_limit = left(p.self_collision_limit);
    for (int i = 0; i < 8; i++) c.self_collision_radii[i] = p.self_collision_radii[i];
    c.workspace_limit_above = left(p.workspace_limit_above); c.workspace_limit_infront = left(p.workspace_limit_infront);
    c.workspace_limit_reach = right(p.workspace_limit_reach); c.workspace_cost_yaw = quad(p.workspace_cost_yaw);
    c.energy_limit_below = left(p.energy_limit_below); c.energy_limit_above = right(p.energy_limit_above);
    c.trajectory_target_scale = p.trajectory_target_scale; c.trajectory_target_maximum = p.trajectory_target_maximum;
    c.trajectory_position_cost = quad(p.trajectory_position_cost); c.trajectory_position_threshold = p.trajectory_position_threshold;
    c.trajectory_velocity_cost = quad(p.trajectory_velocity_cost); c.trajectory_velocity_minimum = p.trajectory_velocity_minimum;
    c.trajectory_velocity_maximum = p.trajectory_velocity_maximum; c.trajectory_velocity_dropoff = p.trajectory_velocity_dropoff;
    c.manipulability_cost = quad(p.manipulability_cost);
    return c;
}

struct Handle {
    std::unique_ptr<TableDynamicsForecast> dynamics_forecast;   // outlives the trajectory's handles
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
        // the reference's own objective (objective/track_point.cpp, compiled unmodified)
        const auto &p = *static_cast<const mppi_b200_track_point *>(params);
        auto d = std::make_unique<RefFrankaDynamics>();
        d->link_mode = p.link_position_mode;
        dyn = std::move(d);
        cost = FrankaRidgeback::TrackPoint::create(track_point_configuration(p));
    } else {
        // the reference's own objective (objective/assisted_manipulation.cpp) reading the forecast wrench through the
        // reference's own DynamicsForecast handle (frankaridgeback/dynamics.cpp)
        const auto &p = *static_cast<const mppi_b200_assisted_manipulation *>(params);
        h->table = std::make_shared<oracle::WrenchTable>();
        FrankaRidgeback::DynamicsForecast::Configuration fc{};
        fc.time_step = c->time_step; fc.horison = c->horison;
        h->dynamics_forecast = std::make_unique<TableDynamicsForecast>(fc, std::make_unique<TableForecast>(h->table), (unsigned)std::ceil(c->horison / c->time_step));
        auto d = std::make_unique<RefFrankaDynamics>();
        d->link_mode = p.link_position_mode;
        d->table = h->table;
        d->forecast = h->dynamics_forecast->create_handle();
        dyn = std::move(d);
        cost = FrankaRidgeback::AssistedManipulation::create(assisted_configuration(p));
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

// ---- the reference's objectives and cost functors evaluated on caller-supplied kinematics ------------------------
// `in` (OBJECTIVE_PROBE_DOUBLES = 159): state[31], end-effector position[3], linear velocity[3], jacobian[6][12]
// row-major, ARM_MOUNT_JOINT position[3], link positions[13][3] (Link enum order), tank energy, has_forecast, wrench[6].
namespace {
struct ProbeDynamics : FrankaRidgeback::Dynamics {
    VectorXd x = VectorXd(31), jq = VectorXd(12), jv = VectorXd(12);
    FrankaRidgeback::EndEffectorState ee;
    Vector3d mount; double links[13][3]; double energy = 0.0;
    std::unique_ptr<FrankaRidgeback::DynamicsForecast::Handle> forecast;
    std::unique_ptr<mppi::Dynamics> copy() override { return nullptr; }
    Eigen::Ref<VectorXd> step(const VectorXd &, double) override { return x; }
    void set_state(const VectorXd &s, double) override { x = s; }
    Eigen::Ref<VectorXd> get_state() override { return x; }
    constexpr int get_control_dof() override { return 12; }
    constexpr int get_state_dof() override { return 31; }
    const VectorXd &get_joint_position() const override { return jq; }
    const VectorXd &get_joint_velocity() const override { return jv; }
    Vector3d get_frame_position(FrankaRidgeback::Frame) override { return mount; }
    Quaterniond get_frame_orientation(FrankaRidgeback::Frame) override { return Quaterniond::Identity(); }
    Vector3d get_link_position(FrankaRidgeback::Link l) override { return Vector3d(links[(int)l][0], links[(int)l][1], links[(int)l][2]); }
    const FrankaRidgeback::EndEffectorState &get_end_effector_state() const override { return ee; }
    double get_joint_power() const override { return 0.0; }
    double get_external_power() const override { return 0.0; }
    double get_tank_energy() const override { return energy; }
    const FrankaRidgeback::DynamicsForecast::Handle *get_forecast() const override { return forecast.get(); }
    Vector6d get_end_effector_simulated_wrench() const override { return Vector6d::Zero(); }
    void add_end_effector_simulated_wrench(Vector6d) override {}
};
}  // namespace

// out[8]: total, joint, self_collision, workspace, energy, velocity, trajectory, manipulability (assisted manipulation);
// out[0] only for track point. One get_cost call per probe after reset (the getters are the logger's running totals).
int ref_objective_probe(int objective, const void *params, const double *in, long count, double *out) {
    auto table = std::make_shared<oracle::WrenchTable>();
    FrankaRidgeback::DynamicsForecast::Configuration fc{};
    fc.time_step = 0.01; fc.horison = 1.0;
    TableDynamicsForecast df(fc, std::make_unique<TableForecast>(table), 100);
    std::unique_ptr<FrankaRidgeback::TrackPoint> tp;
    std::unique_ptr<FrankaRidgeback::AssistedManipulation> am;
    if (objective == MPPI_B200_OBJECTIVE_TRACK_POINT) tp = FrankaRidgeback::TrackPoint::create(track_point_configuration(*static_cast<const mppi_b200_track_point *>(params)));
    else am = FrankaRidgeback::AssistedManipulation::create(assisted_configuration(*static_cast<const mppi_b200_assisted_manipulation *>(params)));
    for (long n = 0; n < count; n++) {
        const double *p = in + n * 159;
        ProbeDynamics d;
        for (int i = 0; i < 31; i++) d.x[i] = p[i];
        for (int i = 0; i < 12; i++) { d.jq[i] = p[i]; d.jv[i] = p[12 + i]; }
        d.ee.position = Vector3d(p[31], p[32], p[33]);
        d.ee.linear_velocity = Vector3d(p[34], p[35], p[36]);
        for (int r = 0; r < 6; r++) for (int j = 0; j < 12; j++) d.ee.jacobian(r, j) = p[37 + r * 12 + j];
        d.mount = Vector3d(p[109], p[110], p[111]);
        for (int l = 0; l < 13; l++) for (int k = 0; k < 3; k++) d.links[l][k] = p[112 + l * 3 + k];
        d.energy = p[151];
        if (p[152] != 0.0) {
            table->present = true; table->t0 = 0.0; table->dt = 0.01;
            table->w.assign(p + 153, p + 159);
            d.forecast = df.create_handle();
        }
        VectorXd control(12);
        double *o = out + n * 8;
        if (tp) {
            o[0] = tp->get_cost(d.x, control, &d, 0.0);
            for (int i = 1; i < 8; i++) o[i] = 0.0;
        } else {
            am->reset(0.0);
            o[0] = am->get_cost(d.x, control, &d, 0.0);
            o[1] = am->get_joint_limit_cost(); o[2] = am->get_self_collision_cost(); o[3] = am->get_workspace_cost();
            o[4] = am->get_energy_tank_cost(); o[5] = am->get_joint_velocity_cost(); o[6] = am->get_trajectory_cost();
            o[7] = am->get_manipulability_cost();
        }
    }
    return 0;
}

// cost.hpp functors: kind 0 QuadraticCost{a=constant,b=linear,c=quadratic}, 1 LeftInverseBarrier{a=bound,b=scale,c=max},
// 2 RightInverseBarrier, 3 UpperLogarithmicBarrier{a=bound,b=scale,c=offset,d=max}, 4 LowerLogarithmicBarrier
void ref_cost_functor(int kind, double a, double b, double c, double d, const double *values, long count, double *out) {
    for (long i = 0; i < count; i++) {
        const double v = values[i];
        switch (kind) {
            case 0: out[i] = QuadraticCost{a, b, c}(v); break;
            case 1: out[i] = LeftInverseBarrierFunction{a, b, c}(v); break;
            case 2: out[i] = RightInverseBarrierFunction{a, b, c}(v); break;
            case 3: out[i] = UpperLogarithmicBarrierFunction{a, b, c, d}(v); break;
            default: out[i] = LowerLogarithmicBarrierFunction{a, b, c, d}(v); break;
        }
    }
}

// energy.hpp (compiled as it lies): n steps of EnergyTank::step from an initial energy; out = energy after each step
void ref_energy_tank(double initial, const double *power, double dt, long count, double *out) {
    EnergyTank tank(initial);
    for (long i = 0; i < count; i++) { tank.step(power[i], dt); out[i] = tank.get_energy(); }
}

// frankaridgeback/state.cpp: make_state(Preset)
void ref_make_state(int preset, double *out) {
    FrankaRidgeback::State s = FrankaRidgeback::make_state((FrankaRidgeback::Preset)preset);
    for (int i = 0; i < 31; i++) out[i] = s[i];
}

// ---- the reference's DynamicsForecast (frankaridgeback/dynamics.cpp:58-138 + dynamics.hpp:122-408) with its own wrench
// forecaster (forecast.cpp / kalman.cpp), rolling the oracle's rigid-body core forward: SURVEY 8f-2. ----
// type: 0 LOCF, 1 AVERAGE, 2 KALMAN (Forecast::Configuration::Type)
void *ref_dynamics_forecast_create(double time_step, double horison, int type, double forecast_horison_or_window, double forecast_time_step, unsigned order) {
    FrankaRidgeback::DynamicsForecast::Configuration c{};
    c.time_step = time_step; c.horison = horison;
    VectorXd zero(6);
    c.end_effector_wrench_forecast.type = (Forecast::Configuration::Type)type;
    if (type == 0) c.end_effector_wrench_forecast.locf = LOCFForecast::Configuration{.observation = zero, .horison = forecast_horison_or_window};
    else if (type == 1) {
        c.end_effector_wrench_forecast.average = AverageForecast::Configuration{.states = 6, .window = forecast_horison_or_window};
        c.end_effector_wrench_forecast.locf = LOCFForecast::Configuration{.observation = zero, .horison = 0.0};   // Forecast::create tests `locf` for AVERAGE (forecast.cpp:20)
    }
    else c.end_effector_wrench_forecast.kalman = KalmanForecast::Configuration{.observed_states = 6, .time_step = forecast_time_step, .horison = forecast_horison_or_window,
                                                                              .order = order, .variance = zero, .initial_state = zero};
    return FrankaRidgeback::DynamicsForecast::create(c, std::make_unique<RefFrankaDynamics>()).release();
}
void ref_dynamics_forecast_destroy(void *h) { delete static_cast<FrankaRidgeback::DynamicsForecast *>(h); }
int ref_dynamics_forecast_steps(void *h) { return (int)static_cast<FrankaRidgeback::DynamicsForecast *>(h)->get_end_effector_trajectory().size(); }
void ref_dynamics_forecast_observe(void *h, const double *wrench, double time) {
    Vector6d w;
    for (int i = 0; i < 6; i++) w[i] = wrench[i];
    static_cast<FrankaRidgeback::DynamicsForecast *>(h)->observe_wrench(w, time);
}
void ref_dynamics_forecast_observe_time(void *h, double time) { static_cast<FrankaRidgeback::DynamicsForecast *>(h)->observe_time(time); }
void ref_dynamics_forecast_run(void *h, const double *state, double time) {
    FrankaRidgeback::State s;
    for (int i = 0; i < 31; i++) s[i] = state[i];
    static_cast<FrankaRidgeback::DynamicsForecast *>(h)->forecast(s, time);
}
// steps x MPPI_B200_DYNAMICS_FORECAST_RECORD, the layout of include/mppi_b200.h
void ref_dynamics_forecast_read(void *h, double *out) {
    auto &f = *static_cast<FrankaRidgeback::DynamicsForecast *>(h);
    const std::size_t steps = f.get_end_effector_trajectory().size();
    for (std::size_t k = 0; k < steps; k++) {
        double *r = out + k * oracle::DF_RECORD;
        const auto &e = f.get_end_effector_trajectory()[k];
        for (int i = 0; i < 12; i++) r[i] = f.get_joint_position()[k][i];
        for (int i = 0; i < 3; i++) { r[12 + i] = e.position[i]; r[19 + i] = e.linear_velocity[i]; r[22 + i] = e.angular_velocity[i]; r[25 + i] = e.linear_acceleration[i]; r[28 + i] = e.angular_acceleration[i]; }
        r[15] = e.orientation.x(); r[16] = e.orientation.y(); r[17] = e.orientation.z(); r[18] = e.orientation.w();
        r[31] = f.get_joint_power_trajectory()[k]; r[32] = f.get_external_power_trajectory()[k]; r[33] = f.get_energy_trajectory()[k];
        for (int i = 0; i < 6; i++) r[34 + i] = f.get_wrench_trajectory()[k][i];
        for (int a = 0; a < 6; a++) for (int j = 0; j < 12; j++) r[40 + a * 12 + j] = e.jacobian(a, j);
    }
}
// parameterise() is protected (dynamics.hpp:361-376): recovered from the record get_end_effector_state(time) returns
long ref_dynamics_forecast_parameterise(void *h, double time) {
    auto &f = *static_cast<FrankaRidgeback::DynamicsForecast *>(h);
    return (long)(&f.get_end_effector_state(time) - f.get_end_effector_trajectory().data());
}
void ref_dynamics_forecast_wrench(void *h, double time, double *out) {
    Vector6d w = static_cast<FrankaRidgeback::DynamicsForecast *>(h)->get_end_effector_wrench(time);
    for (int i = 0; i < 6; i++) out[i] = w[i];
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
