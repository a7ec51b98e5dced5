// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header) for SURVEY §8f-2:
//   FrankaRidgeback::DynamicsForecast::forecast   reference src/frankaridgeback/dynamics.cpp:104-138
//   over PinocchioDynamics::set_state / step      src/frankaridgeback/pinocchio_dynamics.cpp:142-260
// The dynamics roll forward under ZERO control; in the Pinocchio backend add_end_effector_simulated_wrench is
// empty (pinocchio_dynamics.hpp:276) and get_joint_power / get_external_power return 0 (:211-223).
// PARITY UNPINNED for the rigid-body arithmetic underneath (pinocchio absent, see robot_oracle.hpp); the
// recording order, indexing and forecast lookups follow the reference line by line.
// apply_wrench = true is an extension (the torque line the reference leaves commented out,
// pinocchio_dynamics.cpp:240, fed with the forecast wrench): tau += J_ee^T w, J_ee from the previous calculate().
#pragma once
#include <cmath>
#include <limits>
#include <vector>

#include "forecast_oracle.hpp"
#include "systems.hpp"

namespace oracle {

constexpr int DF_RECORD = 112;  // q[12] p[3] quat xyzw[4] v[3] w[3] a[3] alpha[3] joint_power external_power energy wrench[6] J[6][12]

// Eigen::Quaterniond(Matrix3d) (Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl<Other,3,3>), coeffs order x y z w
inline void rotation_to_quaternion(const M3<double> &m, double *xyzw) {
    auto at = [&](int r, int c) { return m.m[3 * r + c]; };
    double t = at(0, 0) + at(1, 1) + at(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        xyzw[3] = 0.5 * t;
        t = 0.5 / t;
        xyzw[0] = (at(2, 1) - at(1, 2)) * t;
        xyzw[1] = (at(0, 2) - at(2, 0)) * t;
        xyzw[2] = (at(1, 0) - at(0, 1)) * t;
    } else {
        int i = 0;
        if (at(1, 1) > at(0, 0)) i = 1;
        if (at(2, 2) > at(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(at(i, i) - at(j, j) - at(k, k) + 1.0);
        xyzw[i] = 0.5 * t;
        t = 0.5 / t;
        xyzw[3] = (at(k, j) - at(j, k)) * t;
        xyzw[j] = (at(j, i) + at(i, j)) * t;
        xyzw[k] = (at(k, i) + at(i, k)) * t;
    }
}

struct DynamicsForecastOracle {
    double time_step, horison;
    unsigned steps;
    bool apply_wrench;
    double last_forecast = std::numeric_limits<double>::min();   // dynamics.cpp:93
    ForecastOracle *wrench_forecast;                              // not owned; nullptr = zero wrench
    RobotCore<double> core;
    std::vector<double> record;                                   // steps x DF_RECORD

    DynamicsForecastOracle(double dt, double h, ForecastOracle *w, bool apply)
        : time_step(dt), horison(h), steps((unsigned)std::ceil(h / dt)), apply_wrench(apply), wrench_forecast(w), record((std::size_t)steps * DF_RECORD, 0.0) {}

    void forecast(const double *state, double time) {  // dynamics.cpp:104-138
        const double control[FR_CONTROL] = {0};
        core.set_state(state, time);
        for (unsigned step = 0; step < steps; ++step) {
            const double t = time + step * time_step;
            double *r = &record[(std::size_t)step * DF_RECORD];
            for (int i = 0; i < FR_NJ; i++) r[i] = core.q[i];
            const auto &ee = core.ee;
            r[12] = ee.position.x; r[13] = ee.position.y; r[14] = ee.position.z;
            rotation_to_quaternion(ee.orientation, r + 15);
            r[19] = ee.linear_velocity.x; r[20] = ee.linear_velocity.y; r[21] = ee.linear_velocity.z;
            r[22] = ee.angular_velocity.x; r[23] = ee.angular_velocity.y; r[24] = ee.angular_velocity.z;
            r[25] = ee.linear_acceleration.x; r[26] = ee.linear_acceleration.y; r[27] = ee.linear_acceleration.z;
            r[28] = ee.angular_acceleration.x; r[29] = ee.angular_acceleration.y; r[30] = ee.angular_acceleration.z;
            r[31] = 0.0; r[32] = 0.0;          // get_joint_power(), get_external_power()
            r[33] = core.tank.energy;
            double w[6] = {0, 0, 0, 0, 0, 0};
            if (wrench_forecast) { const std::vector<double> f = wrench_forecast->forecast(t); for (int k = 0; k < 6; k++) w[k] = f[(std::size_t)k]; }
            for (int k = 0; k < 6; k++) r[34 + k] = w[k];
            for (int a = 0; a < 6; a++) for (int j = 0; j < FR_NJ; j++) r[40 + a * FR_NJ + j] = ee.jacobian[a][j];
            if (apply_wrench) step_with_wrench(control, w);
            else core.step(control, time_step);
        }
        last_forecast = time;
    }

    // RobotCore::step with tau += J^T w between the control torque and calculate() (pinocchio_dynamics.cpp:238-242)
    void step_with_wrench(const double *u, const double *w) {
        auto &c = core;
        const double yaw = c.q[2], cs = std::cos(yaw), sn = std::sin(yaw);
        const double v0 = cs * u[0] - sn * u[1], v1 = sn * u[0] + cs * u[1];
        c.v[0] = v0; c.v[1] = v1; c.v[2] = u[2];
        for (int i = 0; i < FR_NJ; i++) c.tau[i] = 0.0;
        for (int i = 0; i < 7; i++) c.tau[3 + i] = u[3 + i];
        for (int j = 0; j < FR_NJ; j++) { double s = 0.0; for (int a = 0; a < 6; a++) s += c.ee.jacobian[a][j] * w[a]; c.tau[j] += s; }
        c.calculate();
        for (int i = 0; i < FR_NJ; i++) c.v[i] += c.acc[i] * time_step;
        for (int i = 0; i < FR_NJ; i++) c.q[i] += c.v[i] * time_step;
        c.power = 0.0;
        for (int i = 0; i < FR_NJ; i++) c.power += c.tau[i] * c.v[i];
        c.tank.step(c.power, time_step);
        c.state[30] = c.tank.energy;
        for (int i = 0; i < FR_NJ; i++) { c.state[i] = c.q[i]; c.state[FR_NJ + i] = c.v[i]; }
        c.time += time_step;
    }

    // dynamics.hpp:361-376
    long parameterise(double time) const {
        if (time < last_forecast) return 0;
        if (time >= horison) return (long)steps - 1;
        return (long)((time - last_forecast) / time_step);
    }
};

}  // namespace oracle
