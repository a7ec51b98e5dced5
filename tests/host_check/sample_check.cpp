// TEST INFRASTRUCTURE — compiles csrc/sample_core.cuh (what one thread of the sampling kernel produces) for the CPU
// and runs it over every quad index the kernel's grid would enumerate, so the index arithmetic (rollout / step / Philox
// block / element, static rollouts, kept rollouts, injected rows, shard offset) is checked without a GPU. The libm
// stand-ins of the device's MUFU arithmetic exist for this check only; the product library has no CPU path.
#include <cstring>
#include <vector>

#include "host_math.h"
#include "sample_core.cuh"

using namespace mppi_b200;

template <class R, class RI>
static long long run(int T, long long k_begin, long long k_count, const double *U, const unsigned char *kept, const double *ldiag, unsigned long long seed,
                     unsigned long long update_index, int noise_source, const void *injected, void *noise_out) {
    constexpr int NU = 12;
    Frame f;
    std::memset(&f, 0, sizeof f);
    f.seed = seed; f.update_index = update_index; f.noise_source = noise_source;
    DeviceState d;
    std::memset(&d, 0, sizeof d);
    d.batch = 1; d.elem_bytes = sizeof(R); d.nu = NU; d.T = T; d.k_begin = k_begin; d.k_count = k_count;
    d.frame = &f; d.U = const_cast<double *>(U); d.kept = const_cast<unsigned char *>(kept); d.injected = injected;
    d.L_is_diagonal = 1;
    for (int i = 0; i < NU; i++) d.Ldiag[i] = ldiag[i];
    const long long quads = k_count * T * (NU / 4);
    // the kernel's grid: 256-thread blocks over the quads (threads past the end return)
    const long long threads = (quads + 255) / 256 * 256;
    long long stored = 0;
    R *noise = static_cast<R *>(noise_out);
    for (long long g = 0; g < threads; g++) {
        if (g >= quads) continue;
        long long kl; int t, b;
        quad_coordinates<NU>(g, quads, T, &kl, &t, &b);
        R v[4];
        if (!sample_quad<R, RI, NU>(d, d.Ldiag, kl, t, b, v)) continue;
        for (int i = 0; i < 4; i++) noise[(size_t)g * 4 + i] = v[i];
        stored++;
    }
    return stored;
}

// the column kernel's enumeration (one thread per column) must store exactly what the quad enumeration stores
template <class R, class RI>
static long long run_columns(int T, long long k_begin, long long k_count, const double *U, const unsigned char *kept, const double *ldiag, unsigned long long seed,
                             unsigned long long update_index, int noise_source, const void *injected, void *noise_out) {
    constexpr int NU = 12;
    Frame f;
    std::memset(&f, 0, sizeof f);
    f.seed = seed; f.update_index = update_index; f.noise_source = noise_source;
    DeviceState d;
    std::memset(&d, 0, sizeof d);
    d.batch = 1; d.elem_bytes = sizeof(R); d.nu = NU; d.T = T; d.k_begin = k_begin; d.k_count = k_count;
    d.frame = &f; d.U = const_cast<double *>(U); d.kept = const_cast<unsigned char *>(kept); d.injected = injected;
    d.L_is_diagonal = 1;
    for (int i = 0; i < NU; i++) d.Ldiag[i] = ldiag[i];
    const long long cols = k_count * T;
    long long stored = 0;
    R *noise = static_cast<R *>(noise_out);
    for (long long g = 0; g < (cols + 255) / 256 * 256; g++) {
        if (g >= cols) continue;
        long long kl; int t;
        column_coordinates(g, cols, T, &kl, &t);
        R v[NU];
        if (!sample_column<R, RI, NU>(d, d.Ldiag, kl, t, v)) continue;
        for (int i = 0; i < NU; i++) noise[(size_t)g * NU + i] = v[i];
        stored += NU / 4;
    }
    return stored;
}

extern "C" {
long long host_sample_columns(int f32, int injected_is_double, int T, long long k_begin, long long k_count, const double *U, const unsigned char *kept,
                              const double *ldiag, unsigned long long seed, unsigned long long update_index, int noise_source, const void *injected, void *noise_out) {
    if (f32) return injected_is_double ? run_columns<float, double>(T, k_begin, k_count, U, kept, ldiag, seed, update_index, noise_source, injected, noise_out)
                                       : run_columns<float, float>(T, k_begin, k_count, U, kept, ldiag, seed, update_index, noise_source, injected, noise_out);
    return run_columns<double, double>(T, k_begin, k_count, U, kept, ldiag, seed, update_index, noise_source, injected, noise_out);
}
// f32: engine arithmetic is float; injected_is_double: the injected rows are doubles whatever the engine arithmetic
long long host_sample_quads(int f32, int injected_is_double, int T, long long k_begin, long long k_count, const double *U, const unsigned char *kept,
                            const double *ldiag, unsigned long long seed, unsigned long long update_index, int noise_source, const void *injected, void *noise_out) {
    if (f32) return injected_is_double ? run<float, double>(T, k_begin, k_count, U, kept, ldiag, seed, update_index, noise_source, injected, noise_out)
                                       : run<float, float>(T, k_begin, k_count, U, kept, ldiag, seed, update_index, noise_source, injected, noise_out);
    return run<double, double>(T, k_begin, k_count, U, kept, ldiag, seed, update_index, noise_source, injected, noise_out);
}
// both branches of the index arithmetic (32-bit when the enumeration fits, 64-bit otherwise) for arbitrary sizes
void host_quad_coordinates(long long g, long long quads, int T, long long *kl, int *t, int *b) { quad_coordinates<12>(g, quads, T, kl, t, b); }
// engine-creation arithmetic (host_math.h)
void host_sg_weights(int m, int n, double *out) { const auto w = host_math::sg_weights(m, n); for (size_t i = 0; i < w.size(); i++) out[i] = w[i]; }
void host_noise_transform(int n, const double *cov, double *L) { const auto l = host_math::noise_transform(n, cov); for (size_t i = 0; i < l.size(); i++) L[i] = l[i]; }
int host_diagonal_noise_transform(int n, const double *cov, double *ldiag) { return host_math::diagonal_noise_transform(n, cov, ldiag) ? 1 : 0; }
// the chunk-major column enumeration of a rollout block that draws its own noise, and its constants
void host_chase_coordinates(int i, int *chunk, int *rollout, int *t) { chase_column_coordinates(i, chunk, rollout, t); }
int host_chase_steps() { return CHASE_STEPS; }
int host_chase_columns() { return CHASE_COLUMNS; }
void host_philox(const unsigned *ctr4, const unsigned *key2, unsigned *out4) {
    const uint4 r = philox4x32_10(make_uint4(ctr4[0], ctr4[1], ctr4[2], ctr4[3]), make_uint2(key2[0], key2[1]));
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
}
