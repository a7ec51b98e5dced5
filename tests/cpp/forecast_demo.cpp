// Drives the facade's forecast classes (mppi_b200/forecast.hpp: the reference's Forecast / LOCFForecast /
// AverageForecast / KalmanForecast API over the CUDA producer) through the sequences of the reference's own
// forecast test (src/test/case/forecast.cpp:23-160) and prints every forecast for the Python test to check.
#include <cstdio>

#include "mppi_b200/systems.hpp"

static VectorXd vec(std::initializer_list<double> v) { VectorXd o((std::ptrdiff_t)v.size()); std::ptrdiff_t i = 0; for (double x : v) o[i++] = x; return o; }
static void show(const char *tag, const VectorXd &v) { std::printf("%s", tag); for (std::ptrdiff_t i = 0; i < v.size(); i++) std::printf(" %.17g", v[i]); std::printf("\n"); }

int main() {
    {   // test_locf_forecast
        LOCFForecast::Configuration c; c.observation = vec({1.0, 2.0, 3.0});
        auto f = LOCFForecast::create(c);
        if (!f) return 3;
        const double samples[3][3] = {{0.1, -0.7, 0.3}, {0.9, 0.2, -0.4}, {-0.5, 0.6, 0.8}};
        for (auto &s : samples) {
            f->update(vec({s[0], s[1], s[2]}), 0.0);
            show("locf", f->forecast(0.0)); show("locf", f->forecast(1.0)); show("locf", f->forecast(2.0));
        }
    }
    {   // test_average_forecast
        auto f = AverageForecast::create(AverageForecast::Configuration{3, 1.0});
        if (!f) return 3;
        show("average", f->forecast(0.0));
        f->update(vec({0, 1.0, 0}), 1.01); show("average", f->forecast(5.0));
        f->update(vec({0, 1.5, 0}), 1.5); show("average", f->forecast(10.0));
        f->update(vec({1.0, 1.0, 1.0}), 3.0); show("average", f->forecast(3.0));
        for (int i = 0; i < 10; i++) f->update(vec({(double)i, (double)i, (double)i}), 4.5 + i * 0.05);
        show("average", f->forecast(3.5));
        f->update(10.0); show("average", f->forecast(10.0));
        if (AverageForecast::create(AverageForecast::Configuration{3, -1.0})) return 4;   // forecast.cpp:44-47
    }
    {   // test_kalman_linear_forecast in spirit: an exact line, first-order filter, through Forecast::create
        Forecast::Configuration c;
        c.type = Forecast::Configuration::KALMAN;
        if (Forecast::create(c)) return 5;   // selected with no configuration provided
        c.kalman = KalmanForecast::Configuration{6, 0.1, 3.0, 1, VectorXd(), VectorXd()};
        auto f = Forecast::create(c);
        if (!f) return 3;
        double t = 0.0;
        for (int i = 0; i < 60; i++) { t = 0.1 * i; f->update(vec({1.0 * t, -2.0 * t, 0.5 * t, 0.0, 3.0 * t, -1.0 * t}), t); }
        show("kalman", f->forecast(t)); show("kalman", f->forecast(t + 1.0)); show("kalman", f->forecast(t + 2.0));
        f->update(t + 0.05);
        show("kalman", f->forecast(t + 0.5));
    }
    {   // DynamicsForecast (dynamics.hpp:122-408) driven like Actor::act (actor.cpp:160-181)
        FrankaRidgeback::DynamicsForecast::Configuration c;
        c.time_step = 0.01; c.horison = 0.3;
        c.end_effector_wrench_forecast.type = Forecast::Configuration::LOCF;
        c.end_effector_wrench_forecast.locf = LOCFForecast::Configuration{vec({0, 0, 0, 0, 0, 0}), 1e9};
        auto forecast = FrankaRidgeback::DynamicsForecast::create(c);
        if (!forecast) return 3;
        VectorXd w = vec({3.0, -1.0, 2.0, 0.1, 0.2, -0.3});
        forecast->observe_wrench(w, 0.02);
        forecast->observe_time(0.03);
        VectorXd x(31);
        const double q[12] = {0.2, 0.2, 0.78539816339744828, 0.0, 0.62831853071795862, 0.0, -1.5707963267948966, 0.0, 2, 0.78539816339744828, 0.025, 0.025};
        for (int i = 0; i < 12; i++) x[i] = q[i];
        x[12 + 4] = 0.3; x[30] = 10.0;
        forecast->forecast(x, 0.05);
        std::printf("dynforecast %zu %.17g %.17g\n", forecast->get_end_effector_trajectory().size(), forecast->get_last_forecast_time(), forecast->get_time_step());
        const auto &first = forecast->get_end_effector_trajectory().front(), &last = forecast->get_end_effector_trajectory().back();
        std::printf("df_first %.17g %.17g %.17g\n", first.position[0], first.position[1], first.position[2]);
        std::printf("df_last %.17g %.17g %.17g\n", last.position[0], last.position[1], last.position[2]);
        std::printf("df_wrench %.17g %.17g %.17g\n", forecast->get_wrench_trajectory()[5][0], forecast->get_wrench_trajectory()[5][1], forecast->get_end_effector_wrench(0.2)[2]);
        std::printf("df_energy %.17g %.17g\n", forecast->get_energy_trajectory().front(), forecast->get_energy_trajectory().back());
        std::printf("df_lookup %.17g %.17g\n", forecast->get_end_effector_state(0.0).position[0], forecast->get_end_effector_state(0.175).position[0]);
        std::printf("df_q4 %.17g %.17g\n", forecast->get_joint_position()[0][4], forecast->get_joint_position()[29][4]);
        // the handle feeds the controller's dynamics (pinocchio_dynamics.hpp:74-77)
        auto dynamics = FrankaRidgeback::PinocchioDynamics::create(FrankaRidgeback::PinocchioDynamics::default_configuration(), forecast->create_handle());
        if (!dynamics || dynamics->get_forecast()->get() != forecast.get()) return 6;
    }
    return 0;
}
