// TEST INFRASTRUCTURE — drives the REFERENCE's own wrench forecasters (src/controller/forecast.cpp,
// kalman.cpp compiled unmodified from /root/reference by oracle/Makefile `ref`, against oracle/ref_shim) so
// that the restatement in forecast_oracle.hpp and the product's forecast.hpp / k_forecast.cu can be pinned
// against outputs of the reference itself. The shim's matrix inverse is Gauss-Jordan, not
// Eigen's PartialPivLU: the Kalman numbers agree with a real Eigen build to rounding, not bit for bit.
#include <cstring>
#include <memory>

#include "controller/forecast.hpp"

extern "C" {

// type: 0 LOCF, 1 AVERAGE, 2 KALMAN (Forecast::Configuration::Type)
void *ref_forecast_create(int type, int states, double horison_or_window, double time_step, unsigned order, const double *initial) {
    std::unique_ptr<Forecast> f;
    VectorXd init(states);
    for (int i = 0; i < states; i++) init[i] = initial ? initial[i] : 0.0;
    if (type == 0) f = LOCFForecast::create(LOCFForecast::Configuration{.observation = init, .horison = horison_or_window});
    else if (type == 1) f = AverageForecast::create(AverageForecast::Configuration{.states = (unsigned)states, .window = horison_or_window});
    else f = KalmanForecast::create(KalmanForecast::Configuration{.observed_states = (unsigned)states, .time_step = time_step, .horison = horison_or_window,
                                                                  .order = order, .variance = VectorXd(states), .initial_state = init});
    return f.release();
}
void ref_forecast_destroy(void *h) { delete static_cast<Forecast *>(h); }
void ref_forecast_update(void *h, const double *m, int n, double time) {
    VectorXd v(n);
    for (int i = 0; i < n; i++) v[i] = m[i];
    static_cast<Forecast *>(h)->update(v, time);
}
void ref_forecast_update_time(void *h, double time) { static_cast<Forecast *>(h)->update(time); }
void ref_forecast_get(void *h, double time, double *out, int n) {
    VectorXd v = static_cast<Forecast *>(h)->forecast(time);
    for (int i = 0; i < n && i < v.size(); i++) out[i] = v[i];
}

}  // extern "C"
