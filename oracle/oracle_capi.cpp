// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header). C entry points for ctypes
// (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference). The product never
// links or loads this library.
#include <cstring>
#include <memory>
#include <string>

#include "systems.hpp"
#include "forecast_oracle.hpp"
#include "dynamics_forecast_oracle.hpp"

using namespace oracle;

namespace {
thread_local std::string g_error;

struct Handle {
    std::unique_ptr<Trajectory> traj;
    std::shared_ptr<WrenchTable> table;
    int system = 0, objective = 0;
    double dt = 0.01;
};
}  // namespace

extern "C" {

const char *oracle_last_error() { return g_error.c_str(); }

void *oracle_create(const mppi_b200_config *c, const void *params, size_t params_size) {
    g_error.clear();
    Configuration cfg;
    cfg.initial_state.assign(c->state_dof, 0.0);
    cfg.rollouts = c->rollouts;
    cfg.keep_best_rollouts = c->keep_best_rollouts;
    cfg.time_step = c->time_step; cfg.horison = c->horison; cfg.gradient_step = c->gradient_step;
    cfg.cost_scale = c->cost_scale; cfg.cost_discount_factor = c->cost_discount_factor;
    if (c->covariance) cfg.covariance.assign(c->covariance, c->covariance + (size_t)c->covariance_rows * c->covariance_cols);
    cfg.control_bound = c->control_bound != 0;
    if (c->control_min) cfg.control_min.assign(c->control_min, c->control_min + c->control_limits_size);
    if (c->control_max) cfg.control_max.assign(c->control_max, c->control_max + c->control_limits_size);
    if (c->control_default) cfg.control_default = std::vector<double>(c->control_default, c->control_default + c->control_dof);
    cfg.smoothing = c->smoothing != 0; cfg.smoothing_window = c->smoothing_window; cfg.smoothing_order = c->smoothing_order;
    cfg.threads = c->threads > 0 ? (unsigned)c->threads : 0;
    if (c->covariance_rows != c->covariance_cols) { g_error = "controller covariance matrix not square"; return nullptr; }

    auto h = std::make_unique<Handle>();
    h->system = c->system; h->objective = c->objective; h->dt = c->time_step;
    std::unique_ptr<Dynamics> dyn; std::unique_ptr<Cost> cost;
    if (c->system == MPPI_B200_SYSTEM_TOY && c->objective == MPPI_B200_OBJECTIVE_TOY) {
        if (params_size != sizeof(mppi_b200_toy_objective)) { g_error = "objective parameter size"; return nullptr; }
        dyn = std::make_unique<ToyDynamics>();
        cost = std::make_unique<ToyCost>(*static_cast<const mppi_b200_toy_objective *>(params));
    } else if (c->system == MPPI_B200_SYSTEM_FRANKA_RIDGEBACK && c->objective == MPPI_B200_OBJECTIVE_TRACK_POINT) {
        if (params_size != sizeof(mppi_b200_track_point)) { g_error = "objective parameter size"; return nullptr; }
        dyn = std::make_unique<FrankaDynamics>();
        cost = std::make_unique<TrackPointCost>(*static_cast<const mppi_b200_track_point *>(params));
    } else if (c->system == MPPI_B200_SYSTEM_FRANKA_RIDGEBACK && c->objective == MPPI_B200_OBJECTIVE_ASSISTED_MANIPULATION) {
        if (params_size != sizeof(mppi_b200_assisted_manipulation)) { g_error = "objective parameter size"; return nullptr; }
        dyn = std::make_unique<FrankaDynamics>();
        h->table = std::make_shared<WrenchTable>();
        cost = std::make_unique<AssistedManipulationCost>(*static_cast<const mppi_b200_assisted_manipulation *>(params), h->table);
    } else {
        g_error = "unsupported (dynamics, cost) pair";
        return nullptr;
    }
    if (c->state_dof != dyn->get_state_dof()) { g_error = "controller dynamics state dof mismatch"; return nullptr; }
    if (c->control_dof != dyn->get_control_dof()) { g_error = "controller dynamics control dof mismatch"; return nullptr; }
    if (c->covariance_rows != c->control_dof) { g_error = "controller sample variance dof != dynamics and cost control dof"; return nullptr; }
    std::string why;
    h->traj = Trajectory::create(cfg, std::move(dyn), std::move(cost), &why);
    if (!h->traj) { g_error = why; return nullptr; }
    return h.release();
}

void oracle_destroy(void *p) { delete static_cast<Handle *>(p); }

int oracle_update(void *p, const double *state, double time, const double *wrench, const double *injected) {
    auto *h = static_cast<Handle *>(p);
    try {
        if (h->table) {
            h->table->present = wrench != nullptr;
            h->table->t0 = time; h->table->dt = h->dt;
            if (wrench) h->table->w.assign(wrench, wrench + (size_t)h->traj->step_count() * 6);
        }
        h->traj->update(state, time, injected);
    } catch (const std::runtime_error &e) {
        g_error = e.what();
        return g_error.find("all nan") != std::string::npos ? MPPI_B200_ERR_ALL_NAN : MPPI_B200_ERR_TIME;
    } catch (const std::exception &e) {
        g_error = e.what();
        return MPPI_B200_ERR_INVALID;
    }
    return 0;
}

void oracle_set_optimal(void *p, const double *U) { static_cast<Handle *>(p)->traj->set_optimal(U); }
int oracle_get(void *p, double *control, double time) { static_cast<Handle *>(p)->traj->get(control, time); return 0; }

int oracle_read(void *p, int what, void *dst, size_t bytes) {
    auto &t = *static_cast<Handle *>(p)->traj;
    double *out = static_cast<double *>(dst);
    size_t K = t.rollout_count(), n = (size_t)t.control_dof() * t.step_count();
    switch (what) {
        case MPPI_B200_READ_OPTIMAL: if (bytes != n * 8) return -1; std::memcpy(out, t.optimal().data(), bytes); return 0;
        case MPPI_B200_READ_COSTS: if (bytes != K * 8) return -1; for (size_t k = 0; k < K; k++) out[k] = t.rollouts()[k].cost; return 0;
        case MPPI_B200_READ_WEIGHTS: if (bytes != K * 8) return -1; std::memcpy(out, t.weights().data(), bytes); return 0;
        case MPPI_B200_READ_GRADIENT: if (bytes != n * 8) return -1; std::memcpy(out, t.gradient().data(), bytes); return 0;
        case MPPI_B200_READ_NOISE: if (bytes != K * n * 8) return -1; for (size_t k = 0; k < K; k++) std::memcpy(out + k * n, t.rollouts()[k].noise.data(), n * 8); return 0;
        case MPPI_B200_READ_MINMAX: if (bytes != 16) return -1; out[0] = t.m_min; out[1] = t.m_max; return 0;
        case MPPI_B200_READ_OPTIMAL_COST: if (bytes != 8) return -1; out[0] = t.optimal_total_cost(); return 0;
        case MPPI_B200_READ_BREAKDOWN: {
            if (bytes != 64) return -1;
            auto *am = dynamic_cast<AssistedManipulationCost *>(&t.optimal_cost_object());
            Breakdown b = am ? am->bd : Breakdown();
            out[0] = b.joint; out[1] = b.self_collision; out[2] = b.workspace; out[3] = b.energy; out[4] = b.velocity; out[5] = b.trajectory; out[6] = b.manipulability;
            out[7] = t.optimal_total_cost();
            return 0;
        }
        case MPPI_B200_READ_KEPT: {
            auto *o = static_cast<int64_t *>(dst);
            size_t k = bytes / 8;
            if (k > t.ordered().size()) return -1;
            for (size_t i = 0; i < k; i++) o[i] = (int64_t)t.ordered()[i];
            return 0;
        }
    }
    return -1;
}

int oracle_query(void *p, int what, int64_t *v) {
    auto &t = *static_cast<Handle *>(p)->traj;
    switch (what) {
        case MPPI_B200_QUERY_STEP_COUNT: *v = t.step_count(); return 0;
        case MPPI_B200_QUERY_ROLLOUT_COUNT: *v = t.rollout_count(); return 0;
        case MPPI_B200_QUERY_LOCAL_BEGIN: *v = 0; return 0;
        case MPPI_B200_QUERY_LOCAL_COUNT: *v = t.rollout_count(); return 0;
        case MPPI_B200_QUERY_UPDATE_COUNT: *v = (int64_t)t.update_count(); return 0;
        case MPPI_B200_QUERY_ARGMIN: *v = (int64_t)t.m_argmin; return 0;
        case MPPI_B200_QUERY_SHIFT_BY: *v = t.shift_by(); return 0;
        case MPPI_B200_QUERY_STATE_DOF: *v = t.state_dof(); return 0;
        case MPPI_B200_QUERY_CONTROL_DOF: *v = t.control_dof(); return 0;
    }
    return -1;
}

// seconds accumulated in sample / rollout / optimise / filter since creation
void oracle_phase_seconds(void *p, double *out4) { std::memcpy(out4, static_cast<Handle *>(p)->traj->phase_seconds, 32); }

// ---- primitives for known-answer / self-consistency tests --------------------------------------
void oracle_sg_weights(int m, int t, int n, int s, double *out) { auto w = sg_compute_weights(m, t, n, s); std::memcpy(out, w.data(), w.size() * 8); }
double oracle_left_barrier(double bound, double scale, double maxc, double v) { return left_inverse_barrier<double>({bound, scale, maxc}, v); }
double oracle_right_barrier(double bound, double scale, double maxc, double v) { return right_inverse_barrier<double>({bound, scale, maxc}, v); }
double oracle_quadratic(double c0, double c1, double c2, double v) { return quadratic_cost<double>({c0, c1, c2}, v); }
double oracle_upper_log_barrier(double bound, double scale, double offset, double maxc, double v) { return upper_log_barrier({bound, scale, offset, maxc}, v); }
double oracle_lower_log_barrier(double bound, double scale, double offset, double maxc, double v) { return lower_log_barrier({bound, scale, offset, maxc}, v); }
// The objectives of systems.hpp on caller-supplied kinematics (same record as ref_objective_probe, oracle/ref_driver.cpp):
// state[31], end-effector position[3], linear velocity[3], jacobian[6][12] row-major, ARM_MOUNT_JOINT position[3],
// link positions[13][3], tank energy, has_forecast, wrench[6] = 159 doubles in; 8 out (total + the seven terms).
int oracle_objective_probe(int objective, const void *params, const double *in, long count, double *out) {
    for (long n = 0; n < count; n++) {
        const double *p = in + n * 159;
        RobotCore<double> core;
        for (int i = 0; i < FR_STATE; i++) core.state[i] = p[i];
        core.ee.position = {p[31], p[32], p[33]};
        core.ee.linear_velocity = {p[34], p[35], p[36]};
        for (int r = 0; r < 6; r++) for (int j = 0; j < FR_NJ; j++) core.ee.jacobian[r][j] = p[37 + r * 12 + j];
        core.data.oMf_mount.p = {p[109], p[110], p[111]};
        for (int l = 0; l < FR_NLINK; l++) core.link_com_world[l] = {p[112 + l * 3], p[113 + l * 3], p[114 + l * 3]};
        core.tank.set_energy(p[151]);
        double *o = out + n * 8;
        for (int i = 0; i < 8; i++) o[i] = 0.0;
        if (objective == MPPI_B200_OBJECTIVE_TRACK_POINT) {
            auto q = *static_cast<const mppi_b200_track_point *>(params);
            q.link_position_mode = MPPI_B200_LINKS_BODY_COM;   // the probe supplies the link positions
            o[0] = track_point_cost<double>(q, core.state, core);
        } else {
            auto q = *static_cast<const mppi_b200_assisted_manipulation *>(params);
            q.link_position_mode = MPPI_B200_LINKS_BODY_COM;
            Breakdown bd;
            o[0] = assisted_manipulation_cost<double>(q, core.state, core, p[152] != 0.0 ? p + 153 : nullptr, &bd);
            o[1] = bd.joint; o[2] = bd.self_collision; o[3] = bd.workspace; o[4] = bd.energy; o[5] = bd.velocity; o[6] = bd.trajectory; o[7] = bd.manipulability;
        }
    }
    return 0;
}
void oracle_tank(double energy, double power, double dt, double *out2) { EnergyTank<double> t; t.set_energy(energy); t.step(power, dt); out2[0] = t.energy; out2[1] = t.state; }

// SG window driver: runs `updates` reset/add/apply rounds exactly like mppi.cpp:424-440 on one channel
void oracle_sg_run(int steps, int window, unsigned order, int updates, const double *t0s, double dt, const double *u /*updates x steps*/, double *out /*updates x steps*/) {
    SgFilter f(steps, 1, window, order);
    for (int n = 0; n < updates; n++) {
        f.reset(t0s[n]);
        std::vector<double> row(u + (size_t)n * steps, u + (size_t)(n + 1) * steps);
        for (int i = 0; i < steps; i++) f.add_measurement(&row[i], t0s[n] + i * dt);
        for (int i = 0; i < steps; i++) f.apply(&row[i], t0s[n] + i * dt);
        std::memcpy(out + (size_t)n * steps, row.data(), steps * 8);
    }
}

void oracle_robot_fk(const double *q, double *ee_pos, double *mount_pos, double *link_com /* 13 x 3 */) {
    RobotCore<double> c;
    double x[FR_STATE] = {0};
    for (int i = 0; i < FR_NJ; i++) x[i] = q[i];
    c.set_state(x, 0.0);
    ee_pos[0] = c.ee.position.x; ee_pos[1] = c.ee.position.y; ee_pos[2] = c.ee.position.z;
    mount_pos[0] = c.data.oMf_mount.p.x; mount_pos[1] = c.data.oMf_mount.p.y; mount_pos[2] = c.data.oMf_mount.p.z;
    for (int l = 0; l < FR_NLINK; l++) { link_com[3 * l] = c.link_com_world[l].x; link_com[3 * l + 1] = c.link_com_world[l].y; link_com[3 * l + 2] = c.link_com_world[l].z; }
}
void oracle_robot_nle(const double *q, const double *v, double *out) { RobotData<double> d; nonlinear_effects(d, q, v); std::memcpy(out, d.nle, sizeof d.nle); }
void oracle_robot_aba(const double *q, const double *v, const double *tau, double *out) { RobotData<double> d; aba(d, q, v, tau); std::memcpy(out, d.ddq, sizeof d.ddq); }
void oracle_robot_crba(const double *q, double *M) { RobotData<double> d; crba(d, q, M); }
// kinematic quantities the costs read after calculate(): ee position(3), linear velocity(3), angular velocity(3), jacobian 6x12 row-major
void oracle_robot_kinematics(const double *q, const double *v, double *pos, double *lin, double *ang, double *J) {
    RobotCore<double> c;
    double x[FR_STATE] = {0};
    for (int i = 0; i < FR_NJ; i++) { x[i] = q[i]; x[FR_NJ + i] = v[i]; }
    c.set_state(x, 0.0);
    pos[0] = c.ee.position.x; pos[1] = c.ee.position.y; pos[2] = c.ee.position.z;
    lin[0] = c.ee.linear_velocity.x; lin[1] = c.ee.linear_velocity.y; lin[2] = c.ee.linear_velocity.z;
    ang[0] = c.ee.angular_velocity.x; ang[1] = c.ee.angular_velocity.y; ang[2] = c.ee.angular_velocity.z;
    for (int r = 0; r < 6; r++) for (int j = 0; j < FR_NJ; j++) J[r * FR_NJ + j] = c.ee.jacobian[r][j];
}
// n consecutive PinocchioDynamics::step calls from state x (31) under controls u (n x 12); out = n x 31 states
void oracle_robot_rollout(const double *x, const double *u, int n, double dt, double *out) {
    RobotCore<double> c;
    c.set_state(x, 0.0);
    for (int s = 0; s < n; s++) { const double *nx = c.step(u + 12 * s, dt); std::memcpy(out + 31 * s, nx, 31 * 8); }
}

// Op count of ONE rollout-step (cost evaluation + dynamics step) with the counting scalar.
// objective: MPPI_B200_OBJECTIVE_TRACK_POINT / ASSISTED_MANIPULATION. out6 = addsub, mul, div, sqrt, transcendental, compare
uint64_t oracle_count_step_flops(int objective, const void *params, uint64_t *out6) {
    RobotCore<Counted> c;
    Counted x[FR_STATE];
    const double q0[12] = {0.2, 0.2, 0.7853981633974483, 0.0, 0.6283185307179586, 0.0, -1.5707963267948966, 0.0, 2, 0.7853981633974483, 0.025, 0.025};
    for (int i = 0; i < 12; i++) { x[i] = Counted(q0[i]); x[12 + i] = Counted(0.01 * (i + 1)); }
    x[30] = Counted(10.0);
    c.set_state(x, 0.0);
    Counted u[12];
    for (int i = 0; i < 12; i++) u[i] = Counted(0.1 * (i + 1));
    double wrench[6] = {10, 0, 0, 0, 0, 0};
    OpCount::reset();
    Counted cost;
    if (objective == MPPI_B200_OBJECTIVE_TRACK_POINT) cost = track_point_cost<Counted>(*static_cast<const mppi_b200_track_point *>(params), c.state, c);
    else cost = assisted_manipulation_cost<Counted>(*static_cast<const mppi_b200_assisted_manipulation *>(params), c.state, c, wrench, nullptr);
    Counted u2[12];
    for (int i = 0; i < 12; i++) u2[i] = u[i] + Counted(0.5);  // u = U + eps (mppi.cpp:319-322)
    Counted total = cost + cost;                               // discount multiply + accumulate (mppi.cpp:325-337)
    (void)total;
    c.step(u2, Counted(0.01));
    if (out6) { out6[0] = OpCount::addsub; out6[1] = OpCount::mul; out6[2] = OpCount::div; out6[3] = OpCount::sqrt_; out6[4] = OpCount::trans; out6[5] = OpCount::cmp; }
    return OpCount::flops();
}

// ---- wrench forecast producer ------------------------------------------------------
// type: 0 LOCF, 1 AVERAGE, 2 KALMAN (Forecast::Configuration::Type, forecast.hpp:391-396)
void *oracle_forecast_create(int type, int states, double horison_or_window, double time_step, unsigned order, const double *initial) {
    std::vector<double> init(states, 0.0);
    if (initial) init.assign(initial, initial + states);
    ForecastOracle *f = nullptr;
    if (type == 0) f = new LocfOracle(init, horison_or_window);
    else if (type == 1) f = new AverageOracle((unsigned)states, horison_or_window);
    else f = new KalmanOracle((unsigned)states, time_step, horison_or_window, order, init);
    return f;
}
void oracle_forecast_destroy(void *h) { delete static_cast<ForecastOracle *>(h); }
void oracle_forecast_update(void *h, const double *m, int n, double time) { static_cast<ForecastOracle *>(h)->update(std::vector<double>(m, m + n), time); }
void oracle_forecast_update_time(void *h, double time) { static_cast<ForecastOracle *>(h)->update(time); }
void oracle_forecast_get(void *h, double time, double *out, int n) {
    const std::vector<double> v = static_cast<ForecastOracle *>(h)->forecast(time);
    for (int i = 0; i < n && i < (int)v.size(); i++) out[i] = v[(size_t)i];
}

// ---- DynamicsForecast::forecast -------------------------------------------------------
void *oracle_dynamics_forecast_create(double time_step, double horison, void *wrench_forecast, int apply_wrench) {
    return new DynamicsForecastOracle(time_step, horison, static_cast<ForecastOracle *>(wrench_forecast), apply_wrench != 0);
}
void oracle_dynamics_forecast_destroy(void *h) { delete static_cast<DynamicsForecastOracle *>(h); }
int oracle_dynamics_forecast_steps(void *h) { return (int)static_cast<DynamicsForecastOracle *>(h)->steps; }
void oracle_dynamics_forecast_run(void *h, const double *state, double time) { static_cast<DynamicsForecastOracle *>(h)->forecast(state, time); }
void oracle_dynamics_forecast_read(void *h, double *out) {
    auto *f = static_cast<DynamicsForecastOracle *>(h);
    std::memcpy(out, f->record.data(), f->record.size() * sizeof(double));
}
long oracle_dynamics_forecast_parameterise(void *h, double time) { return static_cast<DynamicsForecastOracle *>(h)->parameterise(time); }

}  // extern "C"
