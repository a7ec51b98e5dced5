// Wrench forecast producer on the device, batched: one block per forecaster.
// Replaces, for the 6-component end-effector wrench, the reference's host forecasters
// (src/controller/forecast.hpp:62-140 LOCF, forecast.cpp:41-128 Average, forecast.cpp:130-367 Kalman,
// kalman.cpp:89-152 filter). Compiled with -fmad=false: every product / sum is a separate IEEE operation in
// the reference's order (k ascending per output element, Gauss-Jordan with partial pivoting for the inverse),
// so the tables are bit-identical with the oracle's.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mppi_b200.h"

namespace {

constexpr int OBS = 6;        // wrench components
constexpr int MAXS = 18;      // states = 6 * (order + 1), order <= 2
constexpr int AVG_CAP = 2048; // measurements an AverageForecast can hold inside its window

struct ForecastState {
    int type, batch, S, steps, order;
    double time_step, horison, window;
    // LOCF
    double *obs;          // [batch][6]
    double *valid_until;  // [batch]
    // Average
    double *avg_time;     // [batch][AVG_CAP]
    double *avg_val;      // [batch][AVG_CAP][6]
    int *avg_count;       // [batch]
    double *avg_last;     // [batch]
    double *average;      // [batch][6]
    int *overflow;        // [1]
    // Kalman
    double *F, *Q, *Rm;   // [S][S] column-major, shared
    double *P;            // [batch][S][S]
    double *x, *x_next;   // [batch][S]
    double *meas;         // [batch][S]
    double *pred;         // [batch][steps+1][S]
    double *last_update;  // [batch]
    // staging / output
    double *in;           // [batch][6] measurements
    double *table;        // [batch][tsteps][6]
};

__device__ __forceinline__ double &at(double *m, int S, int i, int j) { return m[i + j * S]; }

// ---- LOCF ------------------------------------------------------------------------------------------
__global__ void k_locf_update(ForecastState f, double time) {
    const int c = blockIdx.x, k = threadIdx.x;
    if (k < OBS) f.obs[c * OBS + k] = f.in[c * OBS + k];
    if (k == 0) f.valid_until[c] = time + f.horison;
}

// ---- Average -----------------------------------------------------------------------------------------
// forecast.cpp:65-122: drop everything not newer than time - window, append, average in insertion order
__global__ void k_average_update(ForecastState f, double time, int has_measurement) {
    const int c = blockIdx.x;
    double *tm = f.avg_time + (size_t)c * AVG_CAP, *val = f.avg_val + (size_t)c * AVG_CAP * OBS;
    __shared__ int s_count, s_skip;
    if (threadIdx.x == 0) {
        int n = f.avg_count[c];
        s_skip = 0;
        if (has_measurement) {
            if (time < f.avg_last[c]) s_skip = 1;   // measurements in the past are ignored
            else {
                f.avg_last[c] = time;
                if (n < AVG_CAP) { tm[n] = time; for (int k = 0; k < OBS; k++) val[n * OBS + k] = f.in[c * OBS + k]; n++; }
                else *f.overflow = 1;
            }
        }
        if (!s_skip && n > 0) {
            // upper_bound(time - window): first element strictly newer; erase everything before it
            int first = 0;
            while (first < n && !(time - f.window < tm[first])) first++;
            if (first > 0) {
                for (int i = first; i < n; i++) { tm[i - first] = tm[i]; for (int k = 0; k < OBS; k++) val[(i - first) * OBS + k] = val[i * OBS + k]; }
                n -= first;
            }
        }
        f.avg_count[c] = n;
        s_count = n;
    }
    __syncthreads();
    if (s_skip) return;
    const int k = threadIdx.x;
    if (k < OBS) {
        const int n = s_count;
        if (n == 0) { f.average[c * OBS + k] = 0.0; return; }
        double total = val[k];
        for (int i = 1; i < n; i++) total += val[i * OBS + k];
        f.average[c * OBS + k] = total / (double)n;
    }
}

// ---- Kalman --------------------------------------------------------------------------------------------
// C = A * B (S x S), one thread per element, k ascending from zero like the reference's products
__device__ __forceinline__ void matmul(const double *A, const double *B, double *C, int S, int tid, int nthreads, bool transpose_b) {
    for (int e = tid; e < S * S; e += nthreads) {
        const int i = e % S, j = e / S;
        double s = 0.0;
        for (int k = 0; k < S; k++) s += A[i + k * S] * (transpose_b ? B[j + k * S] : B[k + j * S]);
        C[e] = s;
    }
}

__global__ void __launch_bounds__(384) k_kalman_update(ForecastState f, double time) {
    extern __shared__ double sm[];
    const int S = f.S, c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    double *P = sm, *A = P + S * S, *Inv = A + S * S, *K = Inv + S * S, *T1 = K + S * S, *z = T1 + S * S, *xs = z + S, *xn = xs + S, *tmp = xn + S;
    __shared__ int s_piv;
    double *gP = f.P + (size_t)c * S * S;
    // measurement vector with finite-difference derivatives (forecast.cpp:292-308)
    if (tid == 0) {
        double *m = f.meas + (size_t)c * S;
        const double dt = time - f.last_update[c];
        double delta[OBS], next[OBS];
        for (int k = 0; k < OBS; k++) delta[k] = (f.in[c * OBS + k] - m[k]) / dt;
        for (int i = 1; i <= f.order; i++) {
            for (int k = 0; k < OBS; k++) next[k] = (delta[k] - m[OBS * i + k]) / dt;
            for (int k = 0; k < OBS; k++) m[OBS * i + k] = delta[k];
            for (int k = 0; k < OBS; k++) delta[k] = next[k];
        }
        for (int k = 0; k < OBS; k++) m[k] = f.in[c * OBS + k];
        f.last_update[c] = time;
    }
    for (int e = tid; e < S * S; e += nt) { P[e] = gP[e]; A[e] = gP[e] + f.Rm[e]; Inv[e] = (e % S == e / S) ? 1.0 : 0.0; }   // H = I: H P H^T + R = P + R
    __syncthreads();
    for (int k = tid; k < S; k += nt) { z[k] = f.meas[(size_t)c * S + k]; xn[k] = f.x_next[(size_t)c * S + k]; }
    // Gauss-Jordan with partial pivoting, the oracle's operation order
    for (int k = 0; k < S; k++) {
        __syncthreads();
        if (tid == 0) {
            int piv = k;
            for (int i = k + 1; i < S; i++) if (fabs(at(A, S, i, k)) > fabs(at(A, S, piv, k))) piv = i;
            s_piv = piv;
        }
        __syncthreads();
        const int piv = s_piv;
        if (piv != k) for (int j = tid; j < S; j += nt) {
            double t = at(A, S, k, j); at(A, S, k, j) = at(A, S, piv, j); at(A, S, piv, j) = t;
            t = at(Inv, S, k, j); at(Inv, S, k, j) = at(Inv, S, piv, j); at(Inv, S, piv, j) = t;
        }
        __syncthreads();
        const double dkk = at(A, S, k, k);
        __syncthreads();
        for (int j = tid; j < S; j += nt) { at(A, S, k, j) /= dkk; at(Inv, S, k, j) /= dkk; }
        __syncthreads();
        // rows i != k: row_i -= f_i * row_k ; the factor column is read before anything is written
        for (int i = tid; i < S; i += nt) tmp[i] = at(A, S, i, k);
        __syncthreads();
        for (int e = tid; e < S * S; e += nt) {
            const int i = e % S, j = e / S;
            const double fi = tmp[i];
            if (i == k || fi == 0.0) continue;
            at(A, S, i, j) -= fi * at(A, S, k, j);
            at(Inv, S, i, j) -= fi * at(Inv, S, k, j);
        }
    }
    __syncthreads();
    matmul(P, Inv, K, S, tid, nt, false);              // K = P H^T (H P H^T + R)^-1
    __syncthreads();
    // x = x_next + K (z - H x_next)
    for (int i = tid; i < S; i += nt) {
        double s = 0.0;
        for (int k = 0; k < S; k++) s += K[i + k * S] * (z[k] - xn[k]);
        xs[i] = xn[i] + s;
    }
    // P = (I - K H) P
    for (int e = tid; e < S * S; e += nt) A[e] = ((e % S == e / S) ? 1.0 : 0.0) - K[e];
    __syncthreads();
    matmul(A, P, T1, S, tid, nt, false);
    __syncthreads();
    // x_next = F x ; P = F P F^T + Q
    for (int i = tid; i < S; i += nt) {
        double s = 0.0;
        for (int k = 0; k < S; k++) s += f.F[i + k * S] * xs[k];
        xn[i] = s;
    }
    matmul(f.F, T1, A, S, tid, nt, false);
    __syncthreads();
    matmul(A, f.F, P, S, tid, nt, true);
    __syncthreads();
    for (int e = tid; e < S * S; e += nt) gP[e] = P[e] + f.Q[e];
    for (int k = tid; k < S; k += nt) { f.x[(size_t)c * S + k] = xs[k]; f.x_next[(size_t)c * S + k] = xn[k]; }
    // predictor: estimation = filter estimate, then `steps` predictions without covariance (forecast.cpp:316-329)
    double *pred = f.pred + (size_t)c * (f.steps + 1) * S;
    for (int k = tid; k < S; k += nt) pred[k] = xs[k];
    __syncthreads();
    // predictor.set_estimation(x): state = x, next = F x (= xn). predict(): state = next; next = F state
    for (int s = 0; s < f.steps; s++) {
        for (int k = tid; k < S; k += nt) { z[k] = xn[k]; pred[(size_t)(s + 1) * S + k] = xn[k]; }
        __syncthreads();
        for (int i = tid; i < S; i += nt) {
            double acc = 0.0;
            for (int k = 0; k < S; k++) acc += f.F[i + k * S] * z[k];
            xn[i] = acc;
        }
        __syncthreads();
    }
}

// Forecast::update(time) of the Kalman forecaster: filter.predict() with covariance (forecast.cpp:332-340)
__global__ void __launch_bounds__(384) k_kalman_predict(ForecastState f, double time) {
    extern __shared__ double sm[];
    const int S = f.S, c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (time <= f.last_update[c]) return;
    double *P = sm, *A = P + S * S, *xs = A + S * S;
    double *gP = f.P + (size_t)c * S * S;
    for (int e = tid; e < S * S; e += nt) P[e] = gP[e];
    for (int k = tid; k < S; k += nt) xs[k] = f.x_next[(size_t)c * S + k];
    __syncthreads();
    for (int i = tid; i < S; i += nt) {
        double s = 0.0;
        for (int k = 0; k < S; k++) s += f.F[i + k * S] * xs[k];
        f.x[(size_t)c * S + i] = xs[i];
        f.x_next[(size_t)c * S + i] = s;
    }
    matmul(f.F, P, A, S, tid, nt, false);
    __syncthreads();
    matmul(A, f.F, P, S, tid, nt, true);
    __syncthreads();
    for (int e = tid; e < S * S; e += nt) gP[e] = P[e] + f.Q[e];
}

// ---- table: forecast(time + k * dt).head(6) for k < tsteps -------------------------------------------------
__global__ void k_table(ForecastState f, double time, double dt, int tsteps) {
    const int c = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tsteps * OBS) return;
    const int k = e / OBS, j = e % OBS;
    const double t_query = time + k * dt;
    double v = 0.0;
    if (f.type == MPPI_B200_FORECAST_LOCF) {
        v = (t_query > f.valid_until[c]) ? 0.0 : f.obs[c * OBS + j];
    } else if (f.type == MPPI_B200_FORECAST_AVERAGE) {
        v = f.average[c * OBS + j];
    } else {
        const double last = f.last_update[c];
        if (!(t_query > last + f.horison)) {
            double t = (t_query - last) / f.time_step;
            int lower = (int)t;
            t -= lower;
            // the reference indexes its table with these unchecked (forecast.cpp:353-366): one column past the
            // end exactly at the horizon (weight 0) and before the start for a query older than the last
            // measurement. Here: the last column, and the estimate itself.
            if (lower < 0) { lower = 0; t = 0.0; }
            if (lower > f.steps) lower = f.steps;
            const int upper = lower + 1 > f.steps ? f.steps : lower + 1;
            const double *pred = f.pred + (size_t)c * (f.steps + 1) * f.S;
            v = (1.0 - t) * pred[(size_t)lower * f.S + j] + t * pred[(size_t)upper * f.S + j];
        }
    }
    f.table[((size_t)c * tsteps + k) * OBS + j] = v;
}

thread_local std::string g_forecast_error;

}  // namespace

struct mppi_b200_forecast {
    mppi_b200_forecast_config cfg{};
    ForecastState s{};
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;
    double *h_in = nullptr;   // pinned staging
    int table_steps = 0;
    std::string error;
};

namespace {
template <class T> T *falloc(mppi_b200_forecast *f, size_t n) {
    void *p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, (n ? n : 1) * sizeof(T));
    f->allocs.push_back(p);
    return static_cast<T *>(p);
}
int ffail(mppi_b200_forecast *f, int code, const std::string &why) { if (f) f->error = why; g_forecast_error = why; return code; }
#define F_TRY(f, call) do { cudaError_t _c = (call); if (_c != cudaSuccess) return ffail((f), MPPI_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_c)); } while (0)
unsigned factorial(unsigned n) { return n <= 1 ? 1 : n * factorial(n - 1); }
}  // namespace

extern "C" {

const char *mppi_b200_forecast_last_error(const mppi_b200_forecast *f) { return f ? f->error.c_str() : g_forecast_error.c_str(); }

void mppi_b200_forecast_destroy(mppi_b200_forecast *f) {
    if (!f) return;
    cudaSetDevice(f->cfg.device);
    if (f->stream) { cudaStreamSynchronize(f->stream); cudaStreamDestroy(f->stream); }
    for (void *p : f->allocs) cudaFree(p);
    if (f->h_in) cudaFreeHost(f->h_in);
    delete f;
}

int mppi_b200_forecast_create(const mppi_b200_forecast_config *c, const double *initial, mppi_b200_forecast **out) {
    if (out) *out = nullptr;
    if (!c || !out) return ffail(nullptr, MPPI_B200_ERR_INVALID, "null argument");
    if (c->batch < 1) return ffail(nullptr, MPPI_B200_ERR_INVALID, "batch");
    if (c->type == MPPI_B200_FORECAST_AVERAGE && c->window < 0.0) return ffail(nullptr, MPPI_B200_ERR_INVALID, "prediction window time is negative");  // forecast.cpp:44-47
    if (c->type == MPPI_B200_FORECAST_KALMAN && (c->order > 2 || !(c->time_step > 0) || !(c->horison > 0))) return ffail(nullptr, MPPI_B200_ERR_INVALID, "kalman order / time_step / horison");
    if (c->type < 0 || c->type > 2) return ffail(nullptr, MPPI_B200_ERR_INVALID, "unknown forecast type");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return ffail(nullptr, MPPI_B200_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    auto *f = new mppi_b200_forecast();
    f->cfg = *c;
    ForecastState &s = f->s;
    const int B = c->batch;
    s.type = c->type; s.batch = B; s.order = (int)c->order; s.time_step = c->time_step; s.horison = c->horison; s.window = c->window;
    s.S = OBS * ((int)c->order + 1);
    s.steps = c->type == MPPI_B200_FORECAST_KALMAN ? (int)std::ceil(c->horison / c->time_step) : 0;
    auto bail = [&](const std::string &why) { mppi_b200_forecast_destroy(f); return ffail(nullptr, MPPI_B200_ERR_CUDA, why); };
    if (cudaSetDevice(c->device) != cudaSuccess) return bail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
    if (cudaMallocHost(&f->h_in, sizeof(double) * B * OBS) != cudaSuccess) return bail("pinned staging");
    bool ok = true;
    auto A = [&](double *&p, size_t n) { p = falloc<double>(f, n); ok = ok && p; };
    A(s.in, (size_t)B * OBS);
    s.overflow = falloc<int>(f, 1); ok = ok && s.overflow;
    std::vector<double> init((size_t)B * OBS, 0.0);
    if (initial) init.assign(initial, initial + (size_t)B * OBS);
    if (c->type == MPPI_B200_FORECAST_LOCF) {
        A(s.obs, (size_t)B * OBS); A(s.valid_until, B);
        if (ok) cudaMemcpy(s.obs, init.data(), init.size() * 8, cudaMemcpyHostToDevice);   // forecast.hpp:126-130: valid_until = 0
    } else if (c->type == MPPI_B200_FORECAST_AVERAGE) {
        A(s.avg_time, (size_t)B * AVG_CAP); A(s.avg_val, (size_t)B * AVG_CAP * OBS); A(s.avg_last, B); A(s.average, (size_t)B * OBS);
        s.avg_count = falloc<int>(f, B); ok = ok && s.avg_count;
    } else {
        const int S = s.S;
        A(s.F, (size_t)S * S); A(s.Q, (size_t)S * S); A(s.Rm, (size_t)S * S); A(s.P, (size_t)B * S * S);
        A(s.x, (size_t)B * S); A(s.x_next, (size_t)B * S); A(s.meas, (size_t)B * S); A(s.pred, (size_t)B * (s.steps + 1) * S); A(s.last_update, B);
        if (ok) {
            // forecast.cpp:212-286: Euler transition matrix, Q = R = P0 = 1e-8 I
            std::vector<double> F((size_t)S * S, 0.0), I8((size_t)S * S, 0.0);
            for (unsigned derivative = 0; derivative <= c->order; derivative++)
                for (unsigned state = 0; state < (unsigned)OBS; state++) {
                    const unsigned row = derivative * OBS + state;
                    for (unsigned i = 0; i <= c->order - derivative; i++) F[row + (size_t)(derivative * OBS + i * OBS + state) * S] = 1.0 / (double)factorial(i) * std::pow(c->time_step, i);
                }
            for (int i = 0; i < S; i++) I8[i + (size_t)i * S] = 1.0 * 1e-8;
            cudaMemcpy(s.F, F.data(), F.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(s.Q, I8.data(), I8.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(s.Rm, I8.data(), I8.size() * 8, cudaMemcpyHostToDevice);
            std::vector<double> P0((size_t)B * S * S, 0.0), x0((size_t)B * S, 0.0), xn((size_t)B * S, 0.0), last(B, -c->time_step);
            for (int b = 0; b < B; b++) {
                for (int i = 0; i < S; i++) P0[(size_t)b * S * S + i + (size_t)i * S] = 1e-8;
                for (int k = 0; k < OBS; k++) x0[(size_t)b * S + k] = init[(size_t)b * OBS + k];
                for (int i = 0; i < S; i++) { double acc = 0.0; for (int k = 0; k < S; k++) acc += F[i + (size_t)k * S] * x0[(size_t)b * S + k]; xn[(size_t)b * S + i] = acc; }
            }
            cudaMemcpy(s.P, P0.data(), P0.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(s.x, x0.data(), x0.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(s.x_next, xn.data(), xn.size() * 8, cudaMemcpyHostToDevice);
            cudaMemcpy(s.last_update, last.data(), last.size() * 8, cudaMemcpyHostToDevice);   // forecast.cpp:196: -time_step
        }
    }
    if (!ok || cudaDeviceSynchronize() != cudaSuccess) return bail("device allocation failed");
    *out = f;
    return MPPI_B200_OK;
}

static size_t kalman_smem(int S) { return sizeof(double) * ((size_t)5 * S * S + 4 * S); }

int mppi_b200_forecast_update(mppi_b200_forecast *f, const double *measurements, double time) {
    if (!f || !measurements) return MPPI_B200_ERR_INVALID;
    F_TRY(f, cudaSetDevice(f->cfg.device));
    F_TRY(f, cudaStreamSynchronize(f->stream));   // the pinned staging buffer is free again
    std::memcpy(f->h_in, measurements, sizeof(double) * f->s.batch * OBS);
    F_TRY(f, cudaMemcpyAsync(f->s.in, f->h_in, sizeof(double) * f->s.batch * OBS, cudaMemcpyHostToDevice, f->stream));
    if (f->s.type == MPPI_B200_FORECAST_LOCF) k_locf_update<<<f->s.batch, 32, 0, f->stream>>>(f->s, time);
    else if (f->s.type == MPPI_B200_FORECAST_AVERAGE) k_average_update<<<f->s.batch, 32, 0, f->stream>>>(f->s, time, 1);
    else k_kalman_update<<<f->s.batch, f->s.S * f->s.S > 384 ? 384 : ((f->s.S * f->s.S + 31) / 32) * 32, kalman_smem(f->s.S), f->stream>>>(f->s, time);
    F_TRY(f, cudaGetLastError());
    return MPPI_B200_OK;
}

int mppi_b200_forecast_update_time(mppi_b200_forecast *f, double time) {
    if (!f) return MPPI_B200_ERR_INVALID;
    F_TRY(f, cudaSetDevice(f->cfg.device));
    if (f->s.type == MPPI_B200_FORECAST_AVERAGE) k_average_update<<<f->s.batch, 32, 0, f->stream>>>(f->s, time, 0);
    else if (f->s.type == MPPI_B200_FORECAST_KALMAN) k_kalman_predict<<<f->s.batch, f->s.S * f->s.S > 384 ? 384 : ((f->s.S * f->s.S + 31) / 32) * 32, kalman_smem(f->s.S), f->stream>>>(f->s, time);
    F_TRY(f, cudaGetLastError());   // LOCFForecast::update(time) is empty (forecast.hpp:108)
    return MPPI_B200_OK;
}

int mppi_b200_forecast_table_device(mppi_b200_forecast *f, double time, double time_step, int32_t steps, const double **device_table) {
    if (!f || steps < 1 || !device_table) return MPPI_B200_ERR_INVALID;
    F_TRY(f, cudaSetDevice(f->cfg.device));
    if (steps > f->table_steps) {
        f->s.table = falloc<double>(f, (size_t)f->s.batch * steps * OBS);
        if (!f->s.table) return ffail(f, MPPI_B200_ERR_CUDA, "device allocation failed");
        f->table_steps = steps;
    }
    k_table<<<dim3((steps * OBS + 127) / 128, f->s.batch), 128, 0, f->stream>>>(f->s, time, time_step, steps);
    F_TRY(f, cudaGetLastError());
    F_TRY(f, cudaStreamSynchronize(f->stream));
    int overflow = 0;
    F_TRY(f, cudaMemcpy(&overflow, f->s.overflow, sizeof(int), cudaMemcpyDeviceToHost));
    if (overflow) return ffail(f, MPPI_B200_ERR_UNSUPPORTED, "average forecast window holds more than 2048 measurements");
    *device_table = f->s.table;
    return MPPI_B200_OK;
}

int mppi_b200_forecast_batch(const mppi_b200_forecast *f) { return f ? f->s.batch : 0; }

int mppi_b200_forecast_table(mppi_b200_forecast *f, double time, double time_step, int32_t steps, double *table) {
    const double *dev = nullptr;
    if (!table) return MPPI_B200_ERR_INVALID;
    const int rc = mppi_b200_forecast_table_device(f, time, time_step, steps, &dev);
    if (rc) return rc;
    F_TRY(f, cudaMemcpy(table, dev, sizeof(double) * f->s.batch * steps * OBS, cudaMemcpyDeviceToHost));
    return MPPI_B200_OK;
}

}  // extern "C"
