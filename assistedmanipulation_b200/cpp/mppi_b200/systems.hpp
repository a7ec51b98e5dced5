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
    std::unique_ptr<mppi::Dynamics> copy() override { auto c = create(m_configuration, m_forecast); if (c) c->m_device_forecast = m_device_forecast; return c; }
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
