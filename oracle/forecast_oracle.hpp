// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header) for SURVEY §8f-1, the wrench-forecast producer
// whose table W[t] the rollout path consumes:
//   LOCFForecast      reference src/controller/forecast.hpp:62-140
//   AverageForecast   src/controller/forecast.cpp:41-128
//   KalmanForecast    src/controller/forecast.cpp:130-367 (update :288-330, forecast :342-367)
//   KalmanFilter      src/controller/kalman.cpp:89-152
// Pinned against the reference's own forecast.cpp / kalman.cpp compiled unmodified (oracle/_ref/
// libforecast_ref.so); matrix products and the inverse use the same operation order as oracle/ref_shim so the
// comparison is bit exact. Eigen's own PartialPivLU inverse would differ in rounding (Eigen is not installed).
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

namespace oracle {

struct Mat {
    long r = 0, c = 0;
    std::vector<double> d;  // column-major
    Mat() = default;
    Mat(long r_, long c_) : r(r_), c(c_), d((std::size_t)r_ * c_, 0.0) {}
    double &operator()(long i, long j) { return d[(std::size_t)(i + j * r)]; }
    double operator()(long i, long j) const { return d[(std::size_t)(i + j * r)]; }
    static Mat identity(long n) { Mat m(n, n); for (long i = 0; i < n; i++) m(i, i) = 1.0; return m; }
};
inline Mat operator*(const Mat &a, const Mat &b) {
    Mat o(a.r, b.c);
    for (long j = 0; j < b.c; j++) for (long k = 0; k < a.c; k++) { const double s = b(k, j); for (long i = 0; i < a.r; i++) o(i, j) += a(i, k) * s; }
    return o;
}
inline Mat operator*(const Mat &a, double s) { Mat o(a); for (auto &x : o.d) x = x * s; return o; }
inline Mat operator+(const Mat &a, const Mat &b) { Mat o(a); for (std::size_t i = 0; i < o.d.size(); i++) o.d[i] = a.d[i] + b.d[i]; return o; }
inline Mat operator-(const Mat &a, const Mat &b) { Mat o(a); for (std::size_t i = 0; i < o.d.size(); i++) o.d[i] = a.d[i] - b.d[i]; return o; }
inline Mat transpose(const Mat &a) { Mat o(a.c, a.r); for (long j = 0; j < a.c; j++) for (long i = 0; i < a.r; i++) o(j, i) = a(i, j); return o; }
inline Mat inverse(const Mat &m) {  // Gauss-Jordan, partial pivoting
    const long n = m.r;
    Mat a(m), inv = Mat::identity(n);
    for (long k = 0; k < n; k++) {
        long piv = k;
        for (long i = k + 1; i < n; i++) if (std::fabs(a(i, k)) > std::fabs(a(piv, k))) piv = i;
        if (piv != k) for (long j = 0; j < n; j++) { std::swap(a(k, j), a(piv, j)); std::swap(inv(k, j), inv(piv, j)); }
        const double dkk = a(k, k);
        for (long j = 0; j < n; j++) { a(k, j) /= dkk; inv(k, j) /= dkk; }
        for (long i = 0; i < n; i++) {
            if (i == k) continue;
            const double f = a(i, k);
            if (f == 0.0) continue;
            for (long j = 0; j < n; j++) { a(i, j) -= f * a(k, j); inv(i, j) -= f * inv(k, j); }
        }
    }
    return inv;
}

struct ForecastOracle {
    virtual ~ForecastOracle() = default;
    virtual void update(const std::vector<double> &measurement, double time) = 0;
    virtual void update(double time) = 0;
    virtual std::vector<double> forecast(double time) = 0;
};

// forecast.hpp:62-140
struct LocfOracle : ForecastOracle {
    double horison, valid_until = 0.0;
    std::vector<double> observation;
    LocfOracle(std::vector<double> obs, double h) : horison(h), observation(std::move(obs)) {}
    void update(const std::vector<double> &m, double time) override { valid_until = time + horison; observation = m; }
    void update(double) override {}
    std::vector<double> forecast(double time) override {
        if (time > valid_until) return std::vector<double>(observation.size(), 0.0);
        return observation;
    }
};

// forecast.cpp:41-128
struct AverageOracle : ForecastOracle {
    double window, last = 0.0;
    std::vector<std::pair<double, std::vector<double>>> buffer;
    std::vector<double> average;
    AverageOracle(unsigned states, double w) : window(w), average(states, 0.0) {}
    void clear_old(double time) {
        if (buffer.empty()) return;
        // upper_bound on (time - window): first element with element.time > time - window
        auto it = std::upper_bound(buffer.begin(), buffer.end(), time - window, [](double t, const std::pair<double, std::vector<double>> &e) { return t < e.first; });
        buffer.erase(buffer.begin(), it);
    }
    void update_average() {
        if (buffer.empty()) { std::fill(average.begin(), average.end(), 0.0); return; }
        std::vector<double> total = buffer[0].second;
        for (std::size_t i = 1; i < buffer.size(); i++) for (std::size_t k = 0; k < total.size(); k++) total[k] += buffer[i].second[k];
        for (std::size_t k = 0; k < total.size(); k++) average[k] = total[k] / (double)buffer.size();
    }
    void update(double time) override { clear_old(time); update_average(); }
    void update(const std::vector<double> &m, double time) override {
        if (time < last) return;
        last = time;
        buffer.emplace_back(time, m);
        clear_old(time);
        update_average();
    }
    std::vector<double> forecast(double) override { return average; }
};

// kalman.cpp:89-152
struct KalmanFilterOracle {
    Mat F, Q, H, R, I, P, x, x_next;
    KalmanFilterOracle(const Mat &F_, const Mat &Q_, const Mat &H_, const Mat &R_, const Mat &x0, const Mat &P0)
        : F(F_), Q(Q_), H(H_), R(R_), I(Mat::identity(F_.r)), P(P0), x(x0), x_next(F_ * x0) {}
    void update(const Mat &z) {
        const Mat K = P * transpose(H) * inverse(H * P * transpose(H) + R);
        x = x_next + K * (z - H * x_next);
        P = (I - K * H) * P;
        x_next = F * x;
        P = F * P * transpose(F) + Q;
    }
    void predict(bool update_covariance = true) {
        x = x_next;
        x_next = F * x;
        if (update_covariance) P = F * P * transpose(F) + Q;
    }
    void set_estimation(const Mat &s) { x = s; x_next = F * s; }
};

// forecast.cpp:130-367 (only meaningful for 6 observed states: the reference hard-codes 6-vectors, :295-308)
struct KalmanOracle : ForecastOracle {
    unsigned observed, order, steps;
    double horison, time_step, last_update;
    Mat measurement, prediction;
    KalmanFilterOracle filter, predictor;

    static unsigned factorial(unsigned n) { return n <= 1 ? 1 : n * factorial(n - 1); }
    static Mat transition(double dt, unsigned obs, unsigned order) {  // forecast.cpp:212-275
        const unsigned states = obs * (order + 1);
        Mat m(states, states);
        for (unsigned derivative = 0; derivative <= order; derivative++)
            for (unsigned state = 0; state < obs; state++) {
                const unsigned row = derivative * obs + state;
                for (unsigned i = 0; i <= order - derivative; i++) m(row, derivative * obs + i * obs + state) = 1.0 / (double)factorial(i) * std::pow(dt, i);
            }
        return m;
    }
    static Mat column(const std::vector<double> &v, unsigned states) { Mat m(states, 1); for (std::size_t i = 0; i < v.size(); i++) m((long)i, 0) = v[i]; return m; }

    KalmanOracle(unsigned obs, double dt, double h, unsigned order_, const std::vector<double> &initial)
        : observed(obs), order(order_), steps((unsigned)std::ceil(h / dt)), horison(h), time_step(dt), last_update(-dt),
          measurement(obs * (order_ + 1), 1), prediction(obs * (order_ + 1), (long)std::ceil(h / dt) + 1),
          filter(transition(dt, obs, order_), Mat::identity(obs * (order_ + 1)) * 1e-8, Mat::identity(obs * (order_ + 1)), Mat::identity(obs * (order_ + 1)) * 1e-8,
                 column(initial, obs * (order_ + 1)), Mat::identity(obs * (order_ + 1)) * 1e-8),
          predictor(filter) {}

    void update(const std::vector<double> &m, double time) override {  // forecast.cpp:288-330
        const double dt = time - last_update;
        double delta[6], next_delta[6];
        for (int k = 0; k < 6; k++) delta[k] = (m[(std::size_t)k] - measurement(k, 0)) / dt;
        for (unsigned i = 1; i <= order; i++) {
            for (int k = 0; k < 6; k++) next_delta[k] = (delta[k] - measurement(6 * i + k, 0)) / dt;
            for (int k = 0; k < 6; k++) measurement(6 * i + k, 0) = delta[k];
            for (int k = 0; k < 6; k++) delta[k] = next_delta[k];
        }
        for (int k = 0; k < 6; k++) measurement(k, 0) = m[(std::size_t)k];
        last_update = time;
        filter.update(measurement);
        predictor.set_estimation(filter.x);
        predictor.P = filter.P;
        for (unsigned k = 0; k < observed * (order + 1); k++) prediction(k, 0) = predictor.x(k, 0);  // the reference keeps the head(observed_states) of the estimate
        for (unsigned i = 0; i < steps; i++) {
            predictor.predict(false);
            for (unsigned k = 0; k < observed * (order + 1); k++) prediction(k, i + 1) = predictor.x(k, 0);
        }
    }
    void update(double time) override { if (time <= last_update) return; filter.predict(); }  // forecast.cpp:332-340
    std::vector<double> forecast(double time) override {  // forecast.cpp:342-367
        if (time > last_update + horison) return std::vector<double>(6, 0.0);
        double t = (time - last_update) / time_step;
        int lower = (int)t;
        t -= lower;
        // unchecked in the reference (one column past the table exactly at the horizon, weight 0; negative for a
        // query older than the last measurement): clamp instead of reading out of bounds
        if (lower < 0) { lower = 0; t = 0.0; }
        if (lower > (int)steps) lower = (int)steps;
        const int upper = std::min(lower + 1, (int)steps);
        std::vector<double> out(6);
        for (int k = 0; k < 6; k++) out[(std::size_t)k] = (1.0 - t) * prediction(k, lower) + t * prediction(k, upper);
        return out;
    }
};

}  // namespace oracle
