// Wrench forecast producer with the reference's API (src/controller/forecast.hpp:14-413), evaluated by the
// batched CUDA producer behind include/mppi_b200.h (csrc/k_forecast.cu). Same class names, Configuration
// fields, create()/update()/forecast() meaning and error behaviour; up to six observed states (the wrench —
// the reference's Kalman forecaster hard-codes six, forecast.cpp:295-308).
//
//   auto forecast = Forecast::create(configuration.wrench_forecast);   // dynamics.cpp: DynamicsForecast::create
//   forecast->update(wrench, time);  forecast->forecast(time + 0.1);
//
// Every call is a device call: there is no host implementation in this library.
#pragma once
#include <iostream>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "mppi_b200.h"
#include "mppi_b200/linalg.hpp"

class Forecast {
public:
    struct Configuration;
    static inline std::unique_ptr<Forecast> create(const Configuration &configuration);
    virtual ~Forecast() { mppi_b200_forecast_destroy(m_handle); }
    Forecast(const Forecast &) = delete;
    Forecast &operator=(const Forecast &) = delete;

    /// forecast.hpp:38
    virtual void update(VectorXd measurement, double time) {
        double m[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < m_states && i < (int)measurement.size(); i++) m[i] = measurement[i];
        check(mppi_b200_forecast_update(m_handle, m, time));
    }
    /// forecast.hpp:47
    virtual void update(double time) { check(mppi_b200_forecast_update_time(m_handle, time)); }
    /// forecast.hpp:55
    virtual VectorXd forecast(double time) {
        double w[6];
        check(mppi_b200_forecast_table(m_handle, time, 1.0, 1, w));
        VectorXd out(m_states);
        for (int i = 0; i < m_states; i++) out[i] = w[i];
        return out;
    }

    /// Not in the reference: the table forecast(time + k * time_step), k < steps, left on the device for
    /// mppi_b200_set_wrench_device (what mppi::Trajectory::update asks for once per update instead of `steps` calls).
    const double *table_device(double time, double time_step, int steps) {
        const double *table = nullptr;
        check(mppi_b200_forecast_table_device(m_handle, time, time_step, steps, &table));
        return table;
    }
    mppi_b200_forecast *handle() { return m_handle; }

protected:
    Forecast(mppi_b200_forecast *handle, int states) : m_handle(handle), m_states(states) {}
    static mppi_b200_forecast *make(int type, int states, double horison, double window, double time_step, unsigned order, const VectorXd *initial) {
        if (states < 1 || states > 6) { std::cerr << "mppi_b200: forecasts hold 1..6 observed states" << std::endl; return nullptr; }
        mppi_b200_forecast_config c{};
        c.type = type; c.batch = 1; c.device = 0; c.order = order; c.time_step = time_step; c.horison = horison; c.window = window;
        double init[6] = {0, 0, 0, 0, 0, 0};
        if (initial) for (int i = 0; i < states && i < (int)initial->size(); i++) init[i] = (*initial)[i];
        mppi_b200_forecast *h = nullptr;
        if (mppi_b200_forecast_create(&c, init, &h) != MPPI_B200_OK) { std::cerr << mppi_b200_forecast_last_error(nullptr) << std::endl; return nullptr; }
        return h;
    }
    void check(int rc) const { if (rc != MPPI_B200_OK) throw std::runtime_error(std::string("mppi_b200 forecast: ") + mppi_b200_forecast_last_error(m_handle)); }
    mppi_b200_forecast *m_handle;
    int m_states;
};

/// forecast.hpp:62-140
class LOCFForecast : public Forecast {
public:
    struct Configuration {
        VectorXd observation;
        double horison = 1e300;   // the reference leaves it to the configuration file; "forever" when unset
    };
    static inline std::unique_ptr<LOCFForecast> create(const Configuration &configuration) {
        mppi_b200_forecast *h = make(MPPI_B200_FORECAST_LOCF, (int)configuration.observation.size(), configuration.horison, 0.0, 0.0, 0, &configuration.observation);
        if (!h) return nullptr;
        return std::unique_ptr<LOCFForecast>(new LOCFForecast(h, (int)configuration.observation.size()));
    }
private:
    using Forecast::Forecast;
};

/// forecast.hpp:147-232
class AverageForecast : public Forecast {
public:
    struct Configuration {
        unsigned int states;
        double window;
    };
    static inline std::unique_ptr<AverageForecast> create(const Configuration &configuration) {
        // "prediction window time is negative" (forecast.cpp:44-47) comes from the library
        mppi_b200_forecast *h = make(MPPI_B200_FORECAST_AVERAGE, (int)configuration.states, 0.0, configuration.window, 0.0, 0, nullptr);
        if (!h) return nullptr;
        return std::unique_ptr<AverageForecast>(new AverageForecast(h, (int)configuration.states));
    }
private:
    using Forecast::Forecast;
};

/// forecast.hpp:239-384
class KalmanForecast : public Forecast {
public:
    struct Configuration {
        unsigned int observed_states;
        double time_step;
        double horison;
        unsigned int order;
        VectorXd variance;        // read by the reference's configuration, unused by its filter (forecast.cpp:130-181)
        VectorXd initial_state;
    };
    static inline std::unique_ptr<KalmanForecast> create(const Configuration &configuration) {
        if (configuration.observed_states != 6) { std::cerr << "mppi_b200: the kalman forecast observes the 6 wrench components" << std::endl; return nullptr; }
        if (configuration.initial_state.size() != 0 && configuration.initial_state.size() != configuration.observed_states) {
            std::cerr << "kalman forecast initial state has wrong size" << std::endl;   // forecast.cpp:138-143 in spirit
            return nullptr;
        }
        mppi_b200_forecast *h = make(MPPI_B200_FORECAST_KALMAN, 6, configuration.horison, 0.0, configuration.time_step, configuration.order,
                                     configuration.initial_state.size() ? &configuration.initial_state : nullptr);
        if (!h) return nullptr;
        return std::unique_ptr<KalmanForecast>(new KalmanForecast(h, 6));
    }
private:
    using Forecast::Forecast;
};

/// forecast.hpp:388-413
struct Forecast::Configuration {
    enum Type { LOCF, AVERAGE, KALMAN } type;
    std::optional<LOCFForecast::Configuration> locf;
    std::optional<AverageForecast::Configuration> average;
    std::optional<KalmanForecast::Configuration> kalman;
};

/// forecast.cpp:6-39
inline std::unique_ptr<Forecast> Forecast::create(const Configuration &configuration) {
    switch (configuration.type) {
        case Configuration::LOCF:
            if (!configuration.locf) { std::cerr << "locf forecast selected with no configuration provided" << std::endl; return nullptr; }
            return LOCFForecast::create(*configuration.locf);
        case Configuration::AVERAGE:
            // the reference tests `locf` here (forecast.cpp:20) and then dereferences `average`: an average forecaster needs
            // both present there; without `average` the reference is undefined, here it is the same refusal
            if (!configuration.locf || !configuration.average) { std::cerr << "average forecast selected with no configuration provided" << std::endl; return nullptr; }
            return AverageForecast::create(*configuration.average);
        case Configuration::KALMAN:
            if (!configuration.kalman) { std::cerr << "kalman forecast selected with no configuration provided" << std::endl; return nullptr; }
            return KalmanForecast::create(*configuration.kalman);
    }
    std::cerr << "unknown forecast type " << (int)configuration.type << "selected" << std::endl;
    return nullptr;
}
