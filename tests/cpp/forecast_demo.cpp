// Drives the facade's forecast classes (mppi_b200/forecast.hpp: the reference's Forecast / LOCFForecast /
// AverageForecast / KalmanForecast API over the CUDA producer) through the sequences of the reference's own
// forecast test (src/test/case/forecast.cpp:23-160) and prints every forecast for the Python test to check.
#include <cstdio>

#include "mppi_b200/forecast.hpp"

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
    return 0;
}
