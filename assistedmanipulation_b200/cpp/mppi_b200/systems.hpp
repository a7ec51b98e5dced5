// The (Dynamics, Cost) pairs a device kernel exists for, with the reference's class names,
// Configuration field names and defaults:
//   FrankaRidgeback::PinocchioDynamics   src/frankaridgeback/pinocchio_dynamics.hpp:30-61
//   FrankaRidgeback::TrackPoint          src/frankaridgeback/objective/track_point.hpp:20-114
//   FrankaRidgeback::AssistedManipulation src/frankaridgeback/objective/assisted_manipulation.hpp:20-206
//   mppi_b200::DoubleIntegrator / PointCost — the toy system of BASELINE.json config 1 (new)
// Their virtual step()/get_cost() exist to satisfy the plug-in interface; the arithmetic lives in the
// rollout kernel, so calling them on the host throws (no CPU fallback).
#pragma once
#include <array>
#include <limits>

#include "mppi_b200/forecast.hpp"
#include "mppi_b200/trajectory.hpp"

namespace mppi_b200 {

[[noreturn]] inline void host_call(const char *what) { throw std::logic_error(std::string(what) + ": evaluated on the device only (mppi_b200 has no CPU path)"); }

class DoubleIntegrator : public mppi::Dynamics, public DeviceBoundDynamics {
public:
    static std::unique_ptr<DoubleIntegrator> create() { return std::unique_ptr<DoubleIntegrator>(new DoubleIntegrator()); }
    std::unique_ptr<mppi::Dynamics> copy() override { return create(); }
    mppi::Ref<VectorXd> step(const VectorXd &, double) override { host_call("DoubleIntegrator::step"); }
    void set_state(const VectorXd &s, double) override { m_state = s; }
    mppi::Ref<VectorXd> get_state() override { return m_state; }
    int get_control_dof() override { return 2; }
    int get_state_dof() override { return 4; }
    int device_system() const override { return MPPI_B200_SYSTEM_TOY; }
private:
    VectorXd m_state = VectorXd(4);
};

class PointCost : public mppi::Cost, public DeviceBoundCost {
public:
    using Configuration = mppi_b200_toy_objective;
    static std::unique_ptr<PointCost> create(const Configuration &c) { return std::unique_ptr<PointCost>(new PointCost(c)); }
    static Configuration default_configuration() { Configuration c; mppi_b200_default_toy_objective(&c); return c; }
    std::unique_ptr<mppi::Cost> copy() override { return create(m_configuration); }
    void reset(double) override {}
    double get_cost(const VectorXd &, const VectorXd &, mppi::Dynamics *, double) override { host_call("PointCost::get_cost"); }
    int get_control_dof() override { return 2; }
    int get_state_dof() override { return 4; }
    int device_objective() const override { return MPPI_B200_OBJECTIVE_TOY; }
    const void *device_params(std::size_t *size) const override { *size = sizeof m_configuration; return &m_configuration; }
private:
    explicit PointCost(const Configuration &c) : m_configuration(c) {}
    Configuration m_configuration;
};

}  // namespace mppi_b200

namespace FrankaRidgeback {

namespace DoF { constexpr std::size_t JOINTS = 12, STATE = 31, CONTROL = 12; }  // dof.hpp:37,63,70

// dynamics.hpp:95-117 (std::array instead of Eigen fixed-size vectors; orientation = quaternion coefficients x y z w)
struct EndEffectorState {
    std::array<double, 3> position{}, linear_velocity{}, angular_velocity{}, linear_acceleration{}, angular_acceleration{};
    std::array<double, 4> orientation{{0.0, 0.0, 0.0, 1.0}};
    std::array<std::array<double, 12>, 6> jacobian{};
};

// dynamics.hpp:122-408: the dynamics rolled forward under zero control with the forecast wrench recorded, one
// trajectory per forecast() call — evaluated on the device (mppi_b200_dynamics_forecast_*, csrc/k_dynforecast.cu).
// The reference hands it the Dynamics instance to roll; here the dynamics are the Pinocchio-backend model the
// kernels are built for, so create() takes only the configuration.
class DynamicsForecast {
public:
    class Handle {   // dynamics.hpp:133-171
    public:
        DynamicsForecast *get() { return m_parent; }
        const DynamicsForecast *get() const { return m_parent; }
        std::unique_ptr<Handle> copy() { return std::unique_ptr<Handle>(new Handle(m_parent)); }
    private:
        friend class DynamicsForecast;
        explicit Handle(DynamicsForecast *parent) : m_parent(parent) {}
        DynamicsForecast *m_parent;
    };
    struct Configuration {
        double time_step;
        double horison;
        Forecast::Configuration end_effector_wrench_forecast;
        bool apply_wrench = false;   // extension: tau += J_ee^T w (pinocchio_dynamics.cpp:240 is commented out in the reference)
    };
    static std::unique_ptr<DynamicsForecast> create(const Configuration &configuration) {
        std::shared_ptr<Forecast> wrench = Forecast::create(configuration.end_effector_wrench_forecast);
        if (!wrench) { std::cerr << "failed to create forecast for end effector wrench" << std::endl; return nullptr; }   // dynamics.cpp:65-68
        mppi_b200_dynamics_forecast_config c{};
        c.batch = 1; c.device = 0; c.time_step = configuration.time_step; c.horison = configuration.horison; c.apply_wrench = configuration.apply_wrench;
        mppi_b200_dynamics_forecast *h = nullptr;
        if (mppi_b200_dynamics_forecast_create(&c, wrench->handle(), &h) != MPPI_B200_OK) { std::cerr << mppi_b200_dynamics_forecast_last_error(nullptr) << std::endl; return nullptr; }
        return std::unique_ptr<DynamicsForecast>(new DynamicsForecast(configuration, std::move(wrench), h));
    }
    ~DynamicsForecast() { mppi_b200_dynamics_forecast_destroy(m_handle); }
    DynamicsForecast(const DynamicsForecast &) = delete;
    DynamicsForecast &operator=(const DynamicsForecast &) = delete;

    std::unique_ptr<Handle> create_handle() { return std::unique_ptr<Handle>(new Handle(this)); }
    void observe_wrench(const VectorXd &wrench, double time) { m_wrench_forecast->update(wrench, time); }   // dynamics.hpp:221-224
    void observe_time(double time) { m_wrench_forecast->update(time); }                                       // dynamics.hpp:231-234

    // dynamics.cpp:104-138
    void forecast(const VectorXd &state, double time) {
        double x[31];
        for (int i = 0; i < 31; i++) x[i] = state[i];
        if (mppi_b200_dynamics_forecast_run(m_handle, x, time) != MPPI_B200_OK) throw std::runtime_error(mppi_b200_dynamics_forecast_last_error(m_handle));
        std::vector<double> rec((std::size_t)m_steps * MPPI_B200_DYNAMICS_FORECAST_RECORD);
        if (mppi_b200_dynamics_forecast_read(m_handle, rec.data(), rec.size() * sizeof(double)) != MPPI_B200_OK) throw std::runtime_error(mppi_b200_dynamics_forecast_last_error(m_handle));
        for (unsigned s = 0; s < m_steps; s++) {
            const double *r = &rec[(std::size_t)s * MPPI_B200_DYNAMICS_FORECAST_RECORD];
            for (int i = 0; i < 12; i++) m_joint_position[s][(std::size_t)i] = r[i];
            EndEffectorState &e = m_end_effector[s];
            for (int i = 0; i < 3; i++) { e.position[i] = r[12 + i]; e.linear_velocity[i] = r[19 + i]; e.angular_velocity[i] = r[22 + i]; e.linear_acceleration[i] = r[25 + i]; e.angular_acceleration[i] = r[28 + i]; }
            for (int i = 0; i < 4; i++) e.orientation[i] = r[15 + i];
            m_joint_power[s] = r[31]; m_external_power[s] = r[32]; m_energy[s] = r[33];
            for (int i = 0; i < 6; i++) m_end_effector_wrench[s][(std::size_t)i] = r[34 + i];
            for (int a = 0; a < 6; a++) for (int j = 0; j < 12; j++) e.jacobian[(std::size_t)a][(std::size_t)j] = r[40 + a * 12 + j];
        }
        m_last_forecast = time;
    }

    double get_last_forecast_time() const { return m_last_forecast; }
    const std::vector<std::array<double, 12>> &get_joint_position() const { return m_joint_position; }
    const EndEffectorState &get_end_effector_state(double time) const { return m_end_effector[(std::size_t)parameterise(time)]; }
    VectorXd get_end_effector_wrench(double time) const { return m_wrench_forecast->forecast(time); }   // dynamics.hpp:275-278
    double get_time_step() const { return m_configuration.time_step; }
    double get_horison() const { return m_configuration.horison; }
    const std::vector<EndEffectorState> &get_end_effector_trajectory() const { return m_end_effector; }
    const std::vector<std::array<double, 6>> &get_wrench_trajectory() const { return m_end_effector_wrench; }
    const std::vector<double> &get_joint_power_trajectory() const { return m_joint_power; }
    const std::vector<double> &get_external_power_trajectory() const { return m_external_power; }
    const std::vector<double> &get_energy_trajectory() const { return m_energy; }
    // the wrench forecast as the device producer the controller reads its table from
    const std::shared_ptr<Forecast> &wrench_forecast() const { return m_wrench_forecast; }

private:
    DynamicsForecast(const Configuration &c, std::shared_ptr<Forecast> wrench, mppi_b200_dynamics_forecast *h)
        : m_configuration(c), m_steps((unsigned)mppi_b200_dynamics_forecast_steps(h)), m_wrench_forecast(std::move(wrench)), m_handle(h),
          m_joint_position(m_steps), m_end_effector(m_steps), m_joint_power(m_steps, 0.0), m_external_power(m_steps, 0.0), m_energy(m_steps, 0.0),
          m_end_effector_wrench(m_steps) {}
    // dynamics.hpp:361-376 (the second test compares the absolute time with the horison, as the reference does)
    std::int64_t parameterise(double time) const {
        if (time < m_last_forecast) return 0;
        if (time >= m_configuration.horison) return (std::int64_t)m_steps - 1;
        return (std::int64_t)((time - m_last_forecast) / m_configuration.time_step);
    }
    Configuration m_configuration;
    const unsigned m_steps;
    double m_last_forecast = std::numeric_limits<double>::min();
    std::shared_ptr<Forecast> m_wrench_forecast;
    mppi_b200_dynamics_forecast *m_handle;
    std::vector<std::array<double, 12>> m_joint_position;
    std::vector<EndEffectorState> m_end_effector;
    std::vector<double> m_joint_power, m_external_power, m_energy;
    std::vector<std::array<double, 6>> m_end_effector_wrench;
};

// pinocchio_dynamics.hpp:30-61. The forecast handle is reduced to what the objective reads from it:
// get_end_effector_wrench(time) (dynamics.hpp:275-278).
class PinocchioDynamics : public mppi::Dynamics, public mppi_b200::DeviceBoundDynamics {
public:
    struct Configuration {
        std::string filename;
        std::string end_effector_frame;
        double energy;
    };
    // pinocchio_dynamics.hpp:56-61
    static Configuration default_configuration() { return Configuration{"", "panda_grasp_joint", 10.0}; }
    using WrenchForecast = std::function<std::array<double, 6>(double time)>;
    static std::unique_ptr<PinocchioDynamics> create() { return create(default_configuration(), WrenchForecast(nullptr)); }
    static std::unique_ptr<PinocchioDynamics> create(Configuration configuration) { return create(configuration, WrenchForecast(nullptr)); }
    static std::unique_ptr<PinocchioDynamics> create(Configuration configuration, WrenchForecast forecast) {
        // the kinematic tree is the one extracted from model/robot.urdf at build time (csrc/robot_model.h)
        if (configuration.end_effector_frame != "panda_grasp_joint") { std::cerr << "mppi_b200: end effector frame must be panda_grasp_joint" << std::endl; return nullptr; }
        return std::unique_ptr<PinocchioDynamics>(new PinocchioDynamics(configuration, std::move(forecast)));
    }
    // The wrench forecast handle as a device producer (mppi_b200/forecast.hpp): its table never leaves the device.
    static std::unique_ptr<PinocchioDynamics> create(Configuration configuration, std::shared_ptr<Forecast> wrench_forecast) {
        auto dynamics = create(configuration, WrenchForecast(nullptr));
        if (dynamics) dynamics->m_device_forecast = std::move(wrench_forecast);
        return dynamics;
    }
    // pinocchio_dynamics.hpp:74-77: the dynamics forecast handle; the controller reads its wrench forecast on the device
    static std::unique_ptr<PinocchioDynamics> create(Configuration configuration, std::unique_ptr<DynamicsForecast::Handle> &&dynamics_forecast_handle) {
        auto dynamics = create(configuration, WrenchForecast(nullptr));
        if (dynamics && dynamics_forecast_handle) { dynamics->m_device_forecast = dynamics_forecast_handle->get()->wrench_forecast(); dynamics->m_handle = std::move(dynamics_forecast_handle); }
        return dynamics;
    }
    DynamicsForecast::Handle *get_forecast() { return m_handle.get(); }   // pinocchio_dynamics.hpp:240-246
    std::unique_ptr<mppi::Dynamics> copy() override { auto c = create(m_configuration, m_forecast); if (c && m_handle) c->m_handle = m_handle->copy(); if (c) c->m_device_forecast = m_device_forecast; return c; }
    mppi::Ref<VectorXd> step(const VectorXd &, double) override { mppi_b200::host_call("PinocchioDynamics::step"); }
    void set_state(const VectorXd &s, double) override { m_state = s; }
    mppi::Ref<VectorXd> get_state() override { return m_state; }
    int get_control_dof() override { return (int)DoF::CONTROL; }
    int get_state_dof() override { return (int)DoF::STATE; }
    int device_system() const override { return MPPI_B200_SYSTEM_FRANKA_RIDGEBACK; }
    bool forecast_wrench(double time, double *w) const override {
        if (!m_forecast) return false;
        const auto v = m_forecast(time);
        for (int i = 0; i < 6; i++) w[i] = v[(std::size_t)i];
        return true;
    }
    const double *forecast_table_device(double time, double time_step, int steps) const override {
        return m_device_forecast ? m_device_forecast->table_device(time, time_step, steps) : nullptr;
    }
private:
    std::shared_ptr<Forecast> m_device_forecast;
    std::unique_ptr<DynamicsForecast::Handle> m_handle;
    PinocchioDynamics(const Configuration &c, WrenchForecast f) : m_configuration(c), m_forecast(std::move(f)) {}
    Configuration m_configuration;
    WrenchForecast m_forecast;
    VectorXd m_state = VectorXd((std::ptrdiff_t)DoF::STATE);
};

// objective/track_point.hpp
class TrackPoint : public mppi::Cost, public mppi_b200::DeviceBoundCost {
public:
    using Configuration = mppi_b200_track_point;  // field names follow track_point.hpp:20-60
    static Configuration default_configuration() { Configuration c; mppi_b200_default_track_point(&c); return c; }
    static std::unique_ptr<TrackPoint> create(const Configuration &c) { return std::unique_ptr<TrackPoint>(new TrackPoint(c)); }
    std::unique_ptr<mppi::Cost> copy() override { return create(m_configuration); }
    void reset(double) override {}
    double get_cost(const VectorXd &, const VectorXd &, mppi::Dynamics *, double) override { mppi_b200::host_call("TrackPoint::get_cost"); }
    int get_control_dof() override { return (int)DoF::CONTROL; }
    int get_state_dof() override { return (int)DoF::STATE; }
    int device_objective() const override { return MPPI_B200_OBJECTIVE_TRACK_POINT; }
    const void *device_params(std::size_t *size) const override { *size = sizeof m_configuration; return &m_configuration; }
private:
    explicit TrackPoint(const Configuration &c) : m_configuration(c) {}
    Configuration m_configuration;
};

// objective/assisted_manipulation.hpp
class AssistedManipulation : public mppi::Cost, public mppi_b200::DeviceBoundCost {
public:
    using Configuration = mppi_b200_assisted_manipulation;  // field names follow assisted_manipulation.hpp:20-99
    static Configuration default_configuration() { Configuration c; mppi_b200_default_assisted_manipulation(&c); return c; }
    static std::unique_ptr<AssistedManipulation> create(const Configuration &c) { return std::unique_ptr<AssistedManipulation>(new AssistedManipulation(c)); }
    std::unique_ptr<mppi::Cost> copy() override { return create(m_configuration); }
    void reset(double) override { for (auto &t : m_terms) t = 0.0; }
    double get_cost(const VectorXd &, const VectorXd &, mppi::Dynamics *, double) override { mppi_b200::host_call("AssistedManipulation::get_cost"); }
    int get_control_dof() override { return (int)DoF::CONTROL; }
    int get_state_dof() override { return (int)DoF::STATE; }
    // assisted_manipulation.hpp:232-262 — totals of the optimal re-rollout, read by the objective logger
    double get_joint_limit_cost() const { return m_terms[0]; }
    double get_self_collision_cost() const { return m_terms[1]; }
    double get_workspace_cost() const { return m_terms[2]; }
    double get_energy_tank_cost() const { return m_terms[3]; }
    double get_joint_velocity_cost() const { return m_terms[4]; }
    double get_trajectory_cost() const { return m_terms[5]; }
    double get_manipulability_cost() const { return m_terms[6]; }
    double get_total_cost() const { return m_terms[7]; }
    int device_objective() const override { return MPPI_B200_OBJECTIVE_ASSISTED_MANIPULATION; }
    const void *device_params(std::size_t *size) const override { *size = sizeof m_configuration; return &m_configuration; }
    void set_optimal_breakdown(const double *t) override { for (int i = 0; i < 8; i++) m_terms[(std::size_t)i] = t[i]; }
private:
    explicit AssistedManipulation(const Configuration &c) : m_configuration(c) {}
    Configuration m_configuration;
    std::array<double, 8> m_terms{};
};

}  // namespace FrankaRidgeback
