// Host side of the C ABI (include/mppi_b200.h): owns the device memory, the streams and the
// per-update orchestration of mppi::Trajectory::update (reference src/controller/mppi.cpp:154-187).
// Nothing here computes the path on the CPU: every stage is a kernel launch (kernels.cuh); the host
// only evaluates the scalar bookkeeping the reference evaluates in double on its caller thread
// (shift_by, mppi.cpp:194; the lerp readout, mppi.cpp:481-512).
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mppi_b200.h"
#include "host_math.h"
#include "kernels.cuh"
#include "model_init.h"
#include "params_convert.h"

using namespace mppi_b200;
using mppi_b200::host_math::noise_transform;
using mppi_b200::host_math::sg_weights;

namespace {

thread_local std::string g_create_error;

// ---- NCCL, resolved at run time so the library loads — and builds — on machines without it: the handful of types and
// enumerators of nccl.h this file uses are declared here (values as in nccl.h 2.x; checked against the loaded library's
// behaviour by the 2-GPU tests) ---------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
typedef int ncclRedOp_t;
constexpr ncclResult_t ncclSuccess = 0;
constexpr ncclDataType_t ncclDouble = 8;          // ncclFloat64
constexpr ncclRedOp_t ncclSum = 0, ncclMax = 2;
struct Nccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string *why) {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { *why = std::string("cannot load libnccl: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather || !CommDestroy) { *why = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
Nccl g_nccl;
std::mutex g_mutex;
bool g_model_uploaded[64] = {false};

}  // namespace

struct mppi_b200_engine {
    mppi_b200_config cfg{};
    DeviceState d{};
    int variant = VAR_TOY;
    bool faithful = false;
    std::vector<unsigned char> params;  // objective block in kernel arithmetic
    // The optimal re-rollout (Trajectory::filter, mppi.cpp:450-479) is evaluated ON DEMAND: every update leaves a snapshot
    // of its inputs and of the sequence it published, and the first read of the optimal cost / breakdown after an update
    // runs the one-thread-per-controller rollout over that snapshot. (Round 1 ran it eagerly on side streams; next to the
    // following update's rollout grid it cost that grid up to 38 % — config 2 FP32: 147 us per update without it, 204 us
    // with it — for a value only the logger reads.)
    static constexpr int SLOTS = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_end = nullptr;
    long long optimal_for = -1;         // update_count the host copy of the optimal cost belongs to
    bool snapshot_valid = false;        // the last update published (its snapshot is complete)
    double *d_opt[SLOTS] = {};          // optimal cost [batch] | breakdown [8 x batch]
    double *h_opt = nullptr;            // pinned, 9 x batch
    // host mirrors (pinned)
    unsigned char *h_frame = nullptr;  // Frame + wrench
    double *h_U = nullptr;             // nu*T (stable copy of the published sequence for get())
    double *h_result = nullptr;        // host-mapped block k_finish writes: U, {-min,max,valid}, argmin, sum w
    double *h_stats = nullptr;         // {-min,max,valid}, argmin bits, sum w, -, optimal cost, breakdown[8]
    size_t frame_bytes = 0;
    // device allocations
    std::vector<void *> allocs;
    unsigned char *d_frame = nullptr;
    double *d_frame_snap[SLOTS] = {};
    double *d_U_snap[SLOTS] = {};
    cudaGraphExec_t graph[SLOTS] = {};
    int graph_launches = 0;
    bool use_graphs = true;
    const double *wrench_device = nullptr;   // forecast table left on the device by the forecast producer (batch x T x 6)
    void *d_zero_row = nullptr;
    void *d_injected = nullptr;
    size_t noise_elems = 0;
    // reference bookkeeping
    int batch = 1;
    double last_shift_time = 0.0, last_rollout_time = 0.0, sg_last_trim = -1.0;
    long long shift_by = 0, update_count = 0, launches = 0;
    bool in_update = false;
    std::vector<double> control_default;
    bool has_default = false;
    std::vector<long long> argmin;       // per controller
    ncclComm_t comm = nullptr;
    // peer-memory exchange (kernels.cuh: PeerExchange): replaces the NCCL calls once mppi_b200_p2p_init has run
    bool p2p = false;
    PeerExchange px{};
    double *mailbox = nullptr;
    int *h_p2p_error = nullptr;          // host-mapped
    std::vector<void *> peer_mappings;
    unsigned long long attempts = 0;
    std::string error;
    std::mutex publish_mutex;            // guards h_U + last_rollout_time: mppi_b200_get may run on another thread during an update (mppi.cpp:178-182,492)
    float last_ms = 0.f;
    bool timing_pending = false;       // last_ms is read from the events on demand (last_device_ms)
    bool profiling = false;
    std::vector<double> weights_total;   // per controller
    std::vector<char> weights_valid;
    cudaEvent_t ev_stage[MPPI_B200_STAGES + 1] = {};
    double stage_s[MPPI_B200_STAGES] = {};
};

namespace {

int fail_create(int code, const std::string &why) { g_create_error = why; return code; }
int fail(mppi_b200_engine *e, int code, const std::string &why) { e->error = why; return code; }
#define STAGE(e, i) do { if ((e)->profiling) cudaEventRecord((e)->ev_stage[i], (e)->stream); } while (0)
#define CUDA_TRY(e, call) do { cudaError_t _c = (call); if (_c != cudaSuccess) return fail((e), MPPI_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_c)); } while (0)

template <class T> T *dev_alloc(mppi_b200_engine *e, size_t count, bool zero = true) {
    void *p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    if (zero) cudaMemset(p, 0, std::max<size_t>(count, 1) * sizeof(T));
    e->allocs.push_back(p);
    return static_cast<T *>(p);
}

template <class R, class CP> void store_params(mppi_b200_engine *e, const CP &cp) {
    auto P = convert<R>(cp);
    e->params.resize(sizeof P);
    std::memcpy(e->params.data(), &P, sizeof P);
}

}  // namespace

extern "C" {

const char *mppi_b200_last_error(const mppi_b200_engine *engine) { return engine ? engine->error.c_str() : g_create_error.c_str(); }

void mppi_b200_destroy(mppi_b200_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (void *p : e->peer_mappings) cudaIpcCloseMemHandle(p);
    if (e->mailbox) cudaFree(e->mailbox);
    if (e->h_p2p_error) cudaFreeHost(e->h_p2p_error);
    for (void *p : e->allocs) cudaFree(p);
    if (e->h_frame) cudaFreeHost(e->h_frame);
    if (e->h_U) cudaFreeHost(e->h_U);
    if (e->h_result) cudaFreeHost(e->h_result);
    if (e->h_stats) cudaFreeHost(e->h_stats);
    for (cudaEvent_t ev : {e->ev_start, e->ev_end}) if (ev) cudaEventDestroy(ev);
    if (e->h_opt) cudaFreeHost(e->h_opt);
    for (cudaGraphExec_t g : e->graph) if (g) cudaGraphExecDestroy(g);
    for (cudaEvent_t ev : e->ev_stage) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int mppi_b200_create(const mppi_b200_config *c, const void *objective_params, size_t objective_params_size, mppi_b200_engine **out) {
    if (out) *out = nullptr;
    if (!c || !out) return fail_create(MPPI_B200_ERR_INVALID, "null argument");
    if (c->abi_version != MPPI_B200_ABI_VERSION) return fail_create(MPPI_B200_ERR_INVALID, "abi version mismatch");
    // the (dynamics, cost) pair must be one a device binding exists for (no CPU fallback)
    int nx = 0, nu = 0;
    if (c->system == MPPI_B200_SYSTEM_TOY && c->objective == MPPI_B200_OBJECTIVE_TOY) { nx = 4; nu = 2; }
    else if (c->system == MPPI_B200_SYSTEM_FRANKA_RIDGEBACK &&
             (c->objective == MPPI_B200_OBJECTIVE_TRACK_POINT || c->objective == MPPI_B200_OBJECTIVE_ASSISTED_MANIPULATION)) { nx = 31; nu = 12; }
    else return fail_create(MPPI_B200_ERR_UNSUPPORTED, "no device implementation for this (dynamics, cost) pair");
    // mppi.cpp:18-69, same order, same messages
    if (c->control_dof != nu) return fail_create(MPPI_B200_ERR_INVALID, "controller dynamics control dof " + std::to_string(c->control_dof) + " != cost control dof " + std::to_string(nu));
    if (c->state_dof != nx) return fail_create(MPPI_B200_ERR_INVALID, "controller dynamics state dof " + std::to_string(c->state_dof) + " != cost state dof " + std::to_string(nx));
    if (c->control_limits_size != nu || !c->control_min || !c->control_max) return fail_create(MPPI_B200_ERR_INVALID, "controller maximum and minimum must have length " + std::to_string(nu));
    if (c->covariance_rows != c->covariance_cols || !c->covariance) return fail_create(MPPI_B200_ERR_INVALID, "controller covariance matrix not square");
    if (c->covariance_rows != nu) return fail_create(MPPI_B200_ERR_INVALID, "controller sample variance dof " + std::to_string(c->covariance_rows) + " != dynamics and cost control dof " + std::to_string(nu));
    if (c->rollouts < 1) return fail_create(MPPI_B200_ERR_INVALID, "trajectory rollouts must be greater than zero");
    if (c->keep_best_rollouts < 0) return fail_create(MPPI_B200_ERR_INVALID, "trajectory cached rollouts cannot be less than zero");
    if (c->threads <= 0) return fail_create(MPPI_B200_ERR_INVALID, "trajectory threads must be positive nonzero");
    if (c->precision != MPPI_B200_FP64 && c->precision != MPPI_B200_FP32) return fail_create(MPPI_B200_ERR_INVALID, "precision");
    if (c->world_size < 1 || c->rank < 0 || c->rank >= c->world_size) return fail_create(MPPI_B200_ERR_INVALID, "rank / world_size");
    if (c->batch < 0) return fail_create(MPPI_B200_ERR_INVALID, "batch");
    if (c->batch > 1 && c->world_size > 1) return fail_create(MPPI_B200_ERR_UNSUPPORTED, "a batched engine cannot also be sharded: give each GPU its own batch");
    if (c->batch > 65535) return fail_create(MPPI_B200_ERR_UNSUPPORTED, "batch exceeds the grid's y extent");
    if (c->world_size > MPPI_MAX_WORLD) return fail_create(MPPI_B200_ERR_UNSUPPORTED, "world_size exceeds " + std::to_string(MPPI_MAX_WORLD));
    // every rank must own at least one rollout: an empty shard would launch empty grids and leave its peers waiting
    if (c->world_size > 1 && c->rollouts + 2 < c->world_size) return fail_create(MPPI_B200_ERR_INVALID, "world_size exceeds the rollout count (rollouts + 2): a rank would own no rollout");
    if (!(c->time_step > 0) || !(c->horison > 0)) return fail_create(MPPI_B200_ERR_INVALID, "time_step and horison must be positive");
    if (c->smoothing && c->smoothing_window > (unsigned)MAX_WINDOW) return fail_create(MPPI_B200_ERR_UNSUPPORTED, "smoothing window too large");
    const int T = (int)std::ceil(c->horison / c->time_step);  // mppi.cpp:85
    const int vec = c->precision == MPPI_B200_FP64 ? 2 : 4;
    if ((nu * T) % vec != 0) return fail_create(MPPI_B200_ERR_UNSUPPORTED, "control_dof * steps must be a multiple of " + std::to_string(vec));
    if (c->system == MPPI_B200_SYSTEM_FRANKA_RIDGEBACK) { std::string why; if (!topology_matches(&why) || !fast_structure_matches(&why)) return fail_create(MPPI_B200_ERR_INVALID, "robot model: " + why); }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail_create(MPPI_B200_ERR_CUDA, "no CUDA device (this engine has no CPU fallback)");
    if (c->device < 0 || c->device >= ndev) return fail_create(MPPI_B200_ERR_INVALID, "device ordinal");

    auto *e = new mppi_b200_engine();
    e->cfg = *c;
    e->cfg.covariance = nullptr; e->cfg.control_min = nullptr; e->cfg.control_max = nullptr; e->cfg.control_default = nullptr;
    e->faithful = c->dynamics_mode == MPPI_B200_DYNAMICS_FAITHFUL;
    if (const char *g = std::getenv("MPPI_B200_NO_GRAPH")) e->use_graphs = !(g[0] == '1');
    if (c->control_default) { e->has_default = true; e->control_default.assign(c->control_default, c->control_default + nu); }
    auto bail = [&](int code, const std::string &why) { mppi_b200_destroy(e); return fail_create(code, why); };
#define CREATE_TRY(call) do { cudaError_t _c = (call); if (_c != cudaSuccess) return bail(MPPI_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_c)); } while (0)
    CREATE_TRY(cudaSetDevice(c->device));
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (!g_model_uploaded[c->device]) { CREATE_TRY(upload_robot_model()); g_model_uploaded[c->device] = true; }
    }
    // objective parameters in kernel arithmetic
    const bool f64 = c->precision == MPPI_B200_FP64;
    if (c->objective == MPPI_B200_OBJECTIVE_TOY) {
        if (objective_params_size != sizeof(mppi_b200_toy_objective) || !objective_params) return bail(MPPI_B200_ERR_INVALID, "objective parameter block size");
        const auto &p = *static_cast<const mppi_b200_toy_objective *>(objective_params);
        e->variant = VAR_TOY;
        if (f64) store_params<double>(e, p); else store_params<float>(e, p);
    } else if (c->objective == MPPI_B200_OBJECTIVE_TRACK_POINT) {
        if (objective_params_size != sizeof(mppi_b200_track_point) || !objective_params) return bail(MPPI_B200_ERR_INVALID, "objective parameter block size");
        const auto &p = *static_cast<const mppi_b200_track_point *>(objective_params);
        e->variant = variant_for(p);
        if (f64) store_params<double>(e, p); else store_params<float>(e, p);
    } else {
        if (objective_params_size != sizeof(mppi_b200_assisted_manipulation) || !objective_params) return bail(MPPI_B200_ERR_INVALID, "objective parameter block size");
        const auto &p = *static_cast<const mppi_b200_assisted_manipulation *>(objective_params);
        e->variant = variant_for(p);
        if (f64) store_params<double>(e, p); else store_params<float>(e, p);
    }

    DeviceState &d = e->d;
    const int B = c->batch > 1 ? c->batch : 1;
    e->batch = B; d.batch = B; d.elem_bytes = f64 ? 8 : 4;
    e->argmin.assign(B, 0); e->weights_total.assign(B, 1.0); e->weights_valid.assign(B, 0);
    d.nu = nu; d.nx = nx; d.T = T;
    d.K_total = c->rollouts + 2;
    // contiguous shards in global index order
    d.k_begin = d.K_total * c->rank / c->world_size;
    d.k_count = d.K_total * (c->rank + 1) / c->world_size - d.k_begin;
    d.keep_best = std::min<long long>(c->keep_best_rollouts, c->rollouts);
    d.dt = c->time_step; d.gradient_step = c->gradient_step; d.cost_scale = c->cost_scale; d.discount = c->cost_discount_factor;
    d.bound = c->control_bound;
    for (int i = 0; i < nu; i++) { d.cmin[i] = c->control_min[i]; d.cmax[i] = c->control_max[i]; }
    const size_t n = (size_t)nu * T, esz = f64 ? 8 : 4;
    e->noise_elems = (size_t)d.k_count * n * B;
    e->frame_bytes = sizeof(Frame) + sizeof(double) * 6 * T;   // per controller

    CREATE_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (cudaEvent_t *ev : {&e->ev_start, &e->ev_end}) CREATE_TRY(cudaEventCreate(ev));
    CREATE_TRY(cudaMallocHost(&e->h_frame, e->frame_bytes * B));
    CREATE_TRY(cudaMallocHost(&e->h_U, n * B * sizeof(double)));
    CREATE_TRY(cudaMallocHost(&e->h_stats, 16 * B * sizeof(double)));
    CREATE_TRY(cudaMallocHost(&e->h_opt, mppi_b200_engine::SLOTS * 9 * B * sizeof(double)));
    std::memset(e->h_opt, 0, mppi_b200_engine::SLOTS * 9 * B * sizeof(double));
    CREATE_TRY(cudaHostAlloc(&e->h_result, (n + 8) * B * sizeof(double), cudaHostAllocMapped));
    std::memset(e->h_result, 0, (n + 8) * B * sizeof(double));
    CREATE_TRY(cudaHostGetDevicePointer((void **)&d.result, e->h_result, 0));
    std::memset(e->h_frame, 0, e->frame_bytes * B);
    std::memset(e->h_U, 0, n * B * sizeof(double));
    std::memset(e->h_stats, 0, 16 * B * sizeof(double));

    bool ok = true;
    // per-controller buffers are laid out [controller][...] (kernels.cuh controller_view)
    auto A = [&](auto *&ptr, size_t count) { using P = std::remove_reference_t<decltype(*ptr)>; ptr = dev_alloc<P>(e, count * B); ok = ok && ptr; };
    auto A1 = [&](auto *&ptr, size_t count) { using P = std::remove_reference_t<decltype(*ptr)>; ptr = dev_alloc<P>(e, count); ok = ok && ptr; };
    e->d_frame = dev_alloc<unsigned char>(e, e->frame_bytes * B); ok = ok && e->d_frame;
    for (int i = 0; i < mppi_b200_engine::SLOTS; i++) {
        e->d_frame_snap[i] = dev_alloc<double>(e, e->frame_bytes / 8 * B); ok = ok && e->d_frame_snap[i];
        e->d_U_snap[i] = dev_alloc<double>(e, n * B); ok = ok && e->d_U_snap[i];
        e->d_opt[i] = dev_alloc<double>(e, 9 * B); ok = ok && e->d_opt[i];
    }
    d.frame_doubles = (int)(e->frame_bytes / 8);
    d.frame_snap = e->d_frame_snap[0]; d.U_snap = e->d_U_snap[0];
    e->d_zero_row = dev_alloc<unsigned char>(e, n * esz); ok = ok && e->d_zero_row;
    d.frame = reinterpret_cast<const Frame *>(e->d_frame);
    d.wrench = reinterpret_cast<const double *>(e->d_frame + sizeof(Frame));
    A(d.U, n); A(d.U_shift, n); A(d.costs, (size_t)d.k_count); A(d.weights, (size_t)d.k_count);
    A(d.kept, (size_t)d.k_count); A(d.kept_list, (size_t)std::max<long long>(d.keep_best, 1));
    d.world = c->world_size; d.rank = c->rank;
    A(d.cand, (size_t)2 * std::max<long long>(d.keep_best, 1)); A(d.cand_all, (size_t)2 * std::max<long long>(d.keep_best, 1) * c->world_size);
    A(d.minmax_enc, 2); A(d.valid_count, 1); A(d.argmin, 1); A(d.finish_count, 1); A(d.minmax, 4); A(d.sums, 1 + n + (size_t)c->world_size);
    A(d.minmax_local, 3 + MPPI_MAX_WORLD); A(d.rollout_done, 1); A(d.reduce_done, 1);
    d.has_px = 0;
    d.weight_blocks = (int)((d.k_count + 255) / 256);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    // few rollouts: fewer partial rows for the second stage; many: two blocks per SM
    d.grad_blocks = (int)std::min<long long>(std::max<long long>(d.k_count / 32, 1), 2LL * sms);
    d.wsum_stride = std::max(d.weight_blocks, d.grad_blocks);
    // One rank and a small rollout set (at most ~5000 rollouts: no more partial rows than k_finish's twelve blocks combine in
    // a microsecond): weights inside the weighted-sum kernel, its partial rows combined by k_finish. Measured, graph replay:
    // config 2 200.6 -> 198.5 us, config 5 877 -> 865 us; with 296 partial rows (K = 16 384) k_finish lost what the two
    // kernels had cost (+12 us), hence the bound. MPPI_B200_FUSED_TAIL=0 / =2: never / whenever the weights fit shared memory.
    static const int fuse_tail = std::getenv("MPPI_B200_FUSED_TAIL") ? std::atoi(std::getenv("MPPI_B200_FUSED_TAIL")) : 1;
    d.fused_tail = (fuse_tail > 0 && c->world_size == 1 && (fuse_tail > 1 || d.grad_blocks <= 160) && (d.k_count + d.grad_blocks - 1) / d.grad_blocks + 4 <= MPPI_FUSED_ROWS) ? 1 : 0;
    A(d.wsum_partial, (size_t)d.wsum_stride); A(d.grad_partial, (size_t)d.grad_blocks * n);
    A(d.gradient, n); A(d.skip, 1); A1(d.L, (size_t)nu * nu); A(d.optimal_cost, 1); A(d.breakdown, 8);
    d.chase = 0;
    A1(d.chase_prepared, 1);
    d.noise = dev_alloc<unsigned char>(e, e->noise_elems * esz); ok = ok && d.noise;
    d.injected = nullptr; d.injected_is_double = 0;
    d.sg_enabled = c->smoothing ? 1 : 0;
    d.sg_window = (int)c->smoothing_window;
    d.sg_len = T + 2 * d.sg_window + 1;
    A(d.sg_uu, (size_t)nu * d.sg_len); A(d.sg_tt, (size_t)nu * d.sg_len); A1(d.sg_weights, (size_t)2 * d.sg_window + 1); A(d.sg_started, (size_t)nu);
    if (!ok) return bail(MPPI_B200_ERR_CUDA, "device allocation failed");

    const std::vector<double> L = noise_transform(nu, c->covariance);
    // diagonal covariance: eps_i = sqrt(Sigma_ii) z_i (host_math.h); otherwise the kernels multiply by L
    d.L_is_diagonal = host_math::diagonal_noise_transform(nu, c->covariance, d.Ldiag) ? 1 : 0;
    CREATE_TRY(cudaMemcpy(d.L, L.data(), L.size() * sizeof(double), cudaMemcpyHostToDevice));
    {   // Noise chase (rollout_core.cuh): the blocks of a rollout grid no larger than the machine draw their own noise on
        // seven more warps each, and the rollouts start on the first steps' noise while the rest is being drawn. One
        // controller, the column sampling path (diagonal covariance, 12 channels), the lean reach-to-pose kernel;
        // MPPI_B200_CHASE=0 keeps the sampling kernel.
        static const bool chase = !(std::getenv("MPPI_B200_CHASE") && std::getenv("MPPI_B200_CHASE")[0] == '0');
        static const bool column_path = !(std::getenv("MPPI_B200_SAMPLE_TILE") && std::atoi(std::getenv("MPPI_B200_SAMPLE_TILE")) != 0);
        if (chase && column_path && B == 1 && nu == 12 && d.L_is_diagonal) {
            int eligible = 0;
            CREATE_TRY(launch_rollout(d, c->precision, e->variant, e->faithful, e->params.data(), false, nullptr, &eligible));
            d.chase = eligible;
        }
    }
    if (c->smoothing) {
        const std::vector<double> w = sg_weights(d.sg_window, (int)c->smoothing_order);
        CREATE_TRY(cudaMemcpy(d.sg_weights, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<double> tt((size_t)nu * d.sg_len * B, -1.0);  // filter.cpp:30-32: times start at -1
        CREATE_TRY(cudaMemcpy(d.sg_tt, tt.data(), tt.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaDeviceSynchronize());
#undef CREATE_TRY
    *out = e;
    g_create_error.clear();
    return MPPI_B200_OK;
}

// ---- one update = host bookkeeping -> [frame upload, K0..K5] -> optimal re-rollout on the side stream -------
namespace {

// mppi.cpp:194-201 evaluated in double on the host exactly as written, plus the pinned frame
int host_prepare(mppi_b200_engine *e, const double *state, double time, const double *wrench, const void *noise, int32_t noise_source, uint64_t seed) {
    DeviceState &d = e->d;
    if (noise_source != MPPI_B200_NOISE_PHILOX && !noise) return fail(e, MPPI_B200_ERR_INVALID, "noise buffer missing");
    const long long shift_by = (long long)((time - e->last_shift_time) / d.dt);
    if (shift_by > d.T) return fail(e, MPPI_B200_ERR_INVALID, "time advanced by more than the horizon (the reference indexes out of range here)");
    if (d.sg_enabled && time < e->sg_last_trim) return fail(e, MPPI_B200_ERR_TIME, "Resetting the window back in the past. Can reset only to larger times than last reset!!!");
    if (shift_by > 0) e->last_shift_time = time;
    e->shift_by = shift_by;
    // the previous update's pinned frames have been consumed: finish waited for its stream work
    for (int c = 0; c < e->batch; c++) {
        unsigned char *base = e->h_frame + (size_t)c * e->frame_bytes;
        Frame *f = reinterpret_cast<Frame *>(base);
        std::memset(f->x0, 0, sizeof f->x0);
        std::memcpy(f->x0, state + (size_t)c * d.nx, sizeof(double) * d.nx);
        f->time = time; f->sg_prev_trim = 0.0; f->shift_by = shift_by; f->seed = seed + (uint64_t)c;
        f->update_index = (unsigned long long)e->update_count;
        f->attempt = e->attempts;
        f->has_wrench = wrench != nullptr || e->wrench_device != nullptr; f->noise_source = noise_source;
        if (wrench && !e->wrench_device) std::memcpy(base + sizeof(Frame), wrench + (size_t)c * 6 * d.T, sizeof(double) * 6 * d.T);
    }
    e->attempts++;
    return MPPI_B200_OK;
}

int enqueue_begin(mppi_b200_engine *e, const void *noise, int32_t noise_source) {
    DeviceState &d = e->d;
    STAGE(e, 0);
    CUDA_TRY(e, cudaMemcpyAsync(e->d_frame, e->h_frame, e->frame_bytes * e->batch, cudaMemcpyHostToDevice, e->stream));
    if (e->wrench_device)   // the forecast producer's table goes straight into the frames' wrench rows
        CUDA_TRY(e, cudaMemcpy2DAsync(e->d_frame + sizeof(Frame), e->frame_bytes, e->wrench_device, sizeof(double) * 6 * d.T, sizeof(double) * 6 * d.T, e->batch, cudaMemcpyDeviceToDevice, e->stream));
    const int prec = e->cfg.precision;
    if (noise_source == MPPI_B200_NOISE_HOST) {
        if (!e->d_injected) { e->d_injected = dev_alloc<double>(e, e->noise_elems, false); if (!e->d_injected) return fail(e, MPPI_B200_ERR_CUDA, "device allocation failed"); }
        // host layout [(K+2)][T][nu] doubles; this engine takes its own shard
        const double *src = static_cast<const double *>(noise) + (size_t)d.k_begin * d.nu * d.T;   // (batched engines are never sharded: k_begin = 0)
        CUDA_TRY(e, cudaMemcpyAsync(e->d_injected, src, e->noise_elems * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        d.injected = e->d_injected; d.injected_is_double = 1;
    } else if (noise_source == MPPI_B200_NOISE_DEVICE) {
        d.injected = noise; d.injected_is_double = 0;  // engine precision, local shard
    } else {
        d.injected = nullptr; d.injected_is_double = 0;
    }
    int launches = 0;
    STAGE(e, 1);
    if (d.keep_best > 0) {
        CUDA_TRY(e, launch_select_kept(d, e->stream)); launches++;
        if (d.world > 1) {
            // warm start over a sharded set: all-gather every rank's best candidates, merge identically everywhere
            if (!e->comm && !e->p2p) return fail(e, MPPI_B200_ERR_UNSUPPORTED, "keep_best_rollouts > 0 on a sharded rollout set needs an in-library exchange (mppi_b200_p2p_init or mppi_b200_comm_init)");
            const long long keep = std::min<long long>(d.keep_best, d.K_total - 2);
            if (e->p2p) {
                CUDA_TRY(e, launch_exchange(d, e->px, EX_CAND, e->stream)); launches++;
            } else {
                ncclResult_t r = g_nccl.AllGather(d.cand, d.cand_all, (size_t)2 * keep, ncclDouble, e->comm, e->stream);
                if (r != ncclSuccess) return fail(e, MPPI_B200_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl all-gather failed");
            }
            CUDA_TRY(e, launch_merge_kept(d, e->stream)); launches++;
        }
    }
    STAGE(e, 2);
    CUDA_TRY(e, launch_sample(d, prec, e->stream, &launches));
    STAGE(e, 3);
    CUDA_TRY(e, launch_rollout(d, prec, e->variant, e->faithful, e->params.data(), false, e->stream)); launches++;
    STAGE(e, 4);
    e->launches += launches;   // (sharded: the rollout grid's last block has published — and with the peer exchange pushed — the min / max payload)
    return MPPI_B200_OK;
}

int enqueue_weights(mppi_b200_engine *e) {
    int launches = e->d.fused_tail ? 0 : 1;
    if (!e->d.fused_tail) CUDA_TRY(e, launch_weights(e->d, e->stream));   // (fused tail: inside the weighted-sum kernel)
    STAGE(e, 5);
    CUDA_TRY(e, launch_gradient(e->d, e->cfg.precision, e->stream, &launches));
    e->launches += launches;
    return MPPI_B200_OK;
}

int enqueue_finish(mppi_b200_engine *e) {
    STAGE(e, 6);
    CUDA_TRY(e, launch_finish(e->d, e->stream));   // writes U and the update's scalars straight into host-mapped memory
    e->launches += 1;
    STAGE(e, 7);
    return MPPI_B200_OK;
}

int enqueue_exchange(mppi_b200_engine *e, int kind) {
    CUDA_TRY(e, launch_exchange(e->d, e->px, kind, e->stream));
    e->launches += 1;
    return MPPI_B200_OK;
}

// Optimal re-rollout (Trajectory::filter, mppi.cpp:450-479), on demand: with no mppi::Filter attached (actor.cpp:100) it
// only produces the optimal cost and its per-term breakdown, which nothing but the logger reads. Runs one thread per
// controller over the snapshot the prepare block of k_sample / k_finish left behind; the engine is idle when this is
// called (mppi_b200_read synchronises first).
int run_optimal(mppi_b200_engine *e) {
    DeviceState &d = e->d;
    DeviceState o = d;
    o.frame = reinterpret_cast<const Frame *>(e->d_frame_snap[0]);
    o.wrench = e->d_frame_snap[0] + sizeof(Frame) / sizeof(double);
    o.U_shift = e->d_U_snap[0];
    o.noise = e->d_zero_row;
    o.optimal_cost = e->d_opt[0]; o.breakdown = e->d_opt[0] + e->batch;   // [batch] | [8 x batch]
    CUDA_TRY(e, launch_rollout(o, e->cfg.precision, e->variant, e->faithful, e->params.data(), true, e->stream));
    e->launches += 1;
    CUDA_TRY(e, cudaMemcpyAsync(e->h_opt, e->d_opt[0], 9 * e->batch * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    e->optimal_for = e->update_count;
    return MPPI_B200_OK;
}

// The update is over for the caller when k_finish's block of results has landed in host memory: its last word carries the
// update's number (finish_publish_stats). Polling that word instead of cudaEventSynchronize(ev_end) takes the kernel's
// retirement, the event's signal and the driver's wake-up out of the caller's latency. The event is still queried now and
// then: a device fault ends the wait with its error instead of a hang (the kernels' own waits time out after 2 s).
int wait_published(mppi_b200_engine *e) {
    const size_t n = (size_t)e->d.nu * e->d.T;
    const double expect = (double)e->attempts;   // host_prepare stamped the frames with attempts - 1
    for (unsigned long long spins = 1;; spins++) {
        bool all = true;
        for (int c = 0; c < e->batch && all; c++) all = *reinterpret_cast<volatile double *>(e->h_result + (size_t)c * (n + 8) + n + 7) == expect;
        if (all) break;
        if ((spins & 255) == 0) {
            const cudaError_t q = cudaEventQuery(e->ev_end);
            if (q == cudaSuccess) break;   // everything the stream was given has run
            if (q != cudaErrorNotReady) return fail(e, MPPI_B200_ERR_CUDA, std::string("update: ") + cudaGetErrorString(q));
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return MPPI_B200_OK;
}

// wait for the results, then the reference's end-of-update bookkeeping (mppi.cpp:178-186)
int host_complete(mppi_b200_engine *e) {
    DeviceState &d = e->d;
    const size_t n = (size_t)d.nu * d.T;
    if (e->profiling) CUDA_TRY(e, cudaEventSynchronize(e->ev_end));
    else if (int rc = wait_published(e)) return rc;
    e->timing_pending = true;
    if (e->profiling) {
        for (int i = 0; i < MPPI_B200_STAGES; i++) { float ms = 0.f; cudaEventElapsedTime(&ms, e->ev_stage[i], e->ev_stage[i + 1]); e->stage_s[i] = ms * 1e-3; }
    }
    e->in_update = false;
    if (e->p2p && *e->h_p2p_error) { *e->h_p2p_error = 0; cudaMemsetAsync(e->px.error_dev, 0, sizeof(int), e->stream); return fail(e, MPPI_B200_ERR_NCCL, "peer exchange timed out: a rank of the sharded rollout set did not arrive"); }
    bool all_nan = false, smoothed = false;
    // the only section a concurrent mppi_b200_get waits for: the copy of the published sequence (the reference holds its
    // mutex around `m_optimal_control = m_optimal_control_shifted` only, mppi.cpp:178-182)
    std::lock_guard<std::mutex> publish(e->publish_mutex);
    for (int c = 0; c < e->batch; c++) {
        const double *res = e->h_result + (size_t)c * (n + 8);
        double *st = e->h_stats + 5 * c;
        std::memcpy(st, res + n, 5 * sizeof(double));   // {-min, max, valid}, argmin, sum w
        // mppi.cpp:368-370: no (or a single) valid rollout is an error; nothing is published for that controller
        if (!(st[2] >= 2.0)) { all_nan = true; continue; }
        std::memcpy(e->h_U + (size_t)c * n, res, n * sizeof(double));
        std::memcpy(&e->argmin[c], st + 3, sizeof(long long));
        const bool early_return = st[1] + st[0] < 1e-6;  // max - min < 1e-6 (mppi.cpp:373-375): weights left untouched
        if (!early_return) { e->weights_total[c] = st[4]; e->weights_valid[c] = 1; smoothed = true; }
    }
    if (all_nan && e->batch == 1) return fail(e, MPPI_B200_ERR_ALL_NAN, "all nan rollouts");
    e->snapshot_valid = e->snapshot_valid || !all_nan;   // an update that threw leaves the previous optimal cost in place (the reference throws before filter())
    if (d.sg_enabled && smoothed) e->sg_last_trim = reinterpret_cast<Frame *>(e->h_frame)->time;
    e->last_rollout_time = reinterpret_cast<Frame *>(e->h_frame)->time;
    e->update_count++;
    if (all_nan) return fail(e, MPPI_B200_ERR_ALL_NAN, "all nan rollouts (in at least one controller of the batch)");
    return MPPI_B200_OK;
}

}  // namespace

int mppi_b200_set_wrench_device(mppi_b200_engine *e, const double *device_table) {
    if (!e) return MPPI_B200_ERR_INVALID;
    if (e->d.nx != 31) return fail(e, MPPI_B200_ERR_UNSUPPORTED, "the toy system has no wrench forecast");
    if (device_table != e->wrench_device) {
        // the copy's source address is part of the captured update: capture again on the next launch
        CUDA_TRY(e, cudaSetDevice(e->cfg.device));
        if (device_table) {   // a table on another GPU would turn the per-update copy into a staged peer copy
            cudaPointerAttributes attr{};
            if (cudaPointerGetAttributes(&attr, device_table) != cudaSuccess || attr.type != cudaMemoryTypeDevice || attr.device != e->cfg.device) {
                cudaGetLastError();
                return fail(e, MPPI_B200_ERR_INVALID, "the wrench table must be device memory of the engine's GPU (create the forecast producer with the same device)");
            }
        }
        CUDA_TRY(e, cudaStreamSynchronize(e->stream));
        for (cudaGraphExec_t &g : e->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
        e->wrench_device = device_table;
    }
    return MPPI_B200_OK;
}

int mppi_b200_update_begin(mppi_b200_engine *e, const double *state, double time, const double *wrench, const void *noise, int32_t noise_source, uint64_t seed) {
    if (!e || !state) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    int rc = host_prepare(e, state, time, wrench, noise, noise_source, seed);
    if (rc) return rc;
    CUDA_TRY(e, cudaEventRecord(e->ev_start, e->stream));
    if ((rc = enqueue_begin(e, noise, noise_source))) return rc;
    e->in_update = true;
    return MPPI_B200_OK;
}

int mppi_b200_update_weights(mppi_b200_engine *e) {
    if (!e || !e->in_update) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    return enqueue_weights(e);
}

int mppi_b200_update_finish(mppi_b200_engine *e) {
    if (!e || !e->in_update) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    int rc = enqueue_finish(e);
    if (rc) return rc;
    CUDA_TRY(e, cudaEventRecord(e->ev_end, e->stream));
    STAGE(e, 8);
    return host_complete(e);
}

int mppi_b200_update_launch(mppi_b200_engine *e, const double *state, double time, const double *wrench, const void *noise, int32_t noise_source, uint64_t seed) {
    if (!e || !state || e->in_update) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    // Production path (in-kernel Philox, one GPU): the whole update is ONE CUDA graph launch. The kernels take
    // no per-update arguments (everything varying lives in the frame), so the graph is captured once per
    // snapshot slot and replayed.
    const bool graph_ok = e->use_graphs && noise_source == MPPI_B200_NOISE_PHILOX && !e->comm && !e->profiling;
    int rc = host_prepare(e, state, time, wrench, noise, noise_source, seed);
    if (rc) return rc;
    const int slot = 0;
    if (graph_ok && !e->graph[slot]) {
        cudaGraph_t g = nullptr;
        const long long before = e->launches;
        CUDA_TRY(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_begin(e, nullptr, MPPI_B200_NOISE_PHILOX);   // (peer exchange: inside the kernels, no launch of its own)
        if (!rc) rc = enqueue_weights(e);
        if (!rc) rc = enqueue_finish(e);
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        e->graph_launches = (int)(e->launches - before);
        e->launches = before;
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return fail(e, MPPI_B200_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&e->graph[slot], g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail(e, MPPI_B200_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
    }
    CUDA_TRY(e, cudaEventRecord(e->ev_start, e->stream));
    if (graph_ok) {
        CUDA_TRY(e, cudaGraphLaunch(e->graph[slot], e->stream));
        e->launches += e->graph_launches;
    } else {
        if ((rc = enqueue_begin(e, noise, noise_source))) return rc;
        if (e->comm && !e->p2p) {
            ncclResult_t r = g_nccl.AllReduce(e->d.minmax_local, e->d.minmax_local, 3 + (size_t)e->d.world, ncclDouble, ncclMax, e->comm, e->stream);
            if (r != ncclSuccess) return fail(e, MPPI_B200_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl all-reduce failed");
        }
        if ((rc = enqueue_weights(e))) return rc;
        if (e->comm && !e->p2p) {
            ncclResult_t r = g_nccl.AllReduce(e->d.sums, e->d.sums, 1 + (size_t)e->d.nu * e->d.T + (size_t)e->d.world, ncclDouble, ncclSum, e->comm, e->stream);
            if (r != ncclSuccess) return fail(e, MPPI_B200_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl all-reduce failed");
        }
        if ((rc = enqueue_finish(e))) return rc;
    }
    CUDA_TRY(e, cudaEventRecord(e->ev_end, e->stream));
    STAGE(e, 8);
    e->in_update = true;
    return MPPI_B200_OK;
}

int mppi_b200_update_wait(mppi_b200_engine *e) {
    if (!e || !e->in_update) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    return host_complete(e);
}

int mppi_b200_update(mppi_b200_engine *e, const double *state, double time, const double *wrench, const void *noise, int32_t noise_source, uint64_t seed) {
    const int rc = mppi_b200_update_launch(e, state, time, wrench, noise, noise_source, seed);
    return rc ? rc : mppi_b200_update_wait(e);
}

int mppi_b200_reduce_buffers(mppi_b200_engine *e, void **minmax, size_t *minmax_count, void **sums, size_t *sums_count) {
    if (!e) return MPPI_B200_ERR_INVALID;
    if (minmax) *minmax = e->d.minmax_local;
    if (minmax_count) *minmax_count = e->d.world > 1 ? 3 + (size_t)e->d.world : 3;
    if (sums) *sums = e->d.sums;
    if (sums_count) *sums_count = 1 + (size_t)e->d.nu * e->d.T + (e->d.world > 1 ? (size_t)e->d.world : 0);
    return MPPI_B200_OK;
}

int mppi_b200_stream(mppi_b200_engine *e, void **cuda_stream) {
    if (!e || !cuda_stream) return MPPI_B200_ERR_INVALID;
    *cuda_stream = e->stream;
    return MPPI_B200_OK;
}

int mppi_b200_synchronize(mppi_b200_engine *e) {
    if (!e) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return MPPI_B200_OK;
}

int mppi_b200_comm_unique_id(void *id128) {
    std::string why;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_nccl.load(&why)) { g_create_error = why; return MPPI_B200_ERR_NCCL; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return MPPI_B200_ERR_NCCL; }
    std::memcpy(id128, &id, 128);
    return MPPI_B200_OK;
}

int mppi_b200_comm_init(mppi_b200_engine *e, const void *id128) {
    if (!e || !id128) return MPPI_B200_ERR_INVALID;
    std::string why;
    { std::lock_guard<std::mutex> lock(g_mutex); if (!g_nccl.load(&why)) return fail(e, MPPI_B200_ERR_NCCL, why); }
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&e->comm, e->cfg.world_size, id, e->cfg.rank);
    if (r != ncclSuccess) return fail(e, MPPI_B200_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclCommInitRank failed");
    return MPPI_B200_OK;
}

// ---- peer-memory exchange set-up ------------------------------------------------------------------------------
static int p2p_allocate(mppi_b200_engine *e) {
    if (e->mailbox) return MPPI_B200_OK;
    const DeviceState &d = e->d;
    if (d.world < 2 || d.world > MPPI_MAX_WORLD) return fail(e, MPPI_B200_ERR_UNSUPPORTED, "peer exchange: world size 2.." + std::to_string(MPPI_MAX_WORLD));
    PeerExchange &px = e->px;
    px.world = d.world; px.rank = d.rank;
    const long long keep = std::max<long long>(1, std::min<long long>(d.keep_best, d.K_total - 2));
    px.count[EX_MINMAX] = 3 + d.world; px.count[EX_SUMS] = 1 + d.nu * d.T + d.world; px.count[EX_CAND] = (int)(2 * keep);
    long long at = 0;
    for (int parity = 0; parity < 2; parity++)
        for (int k = 0; k < EX_KINDS; k++) { px.offset[parity][k] = at; at += (long long)d.world * ((px.count[k] + 1) & ~1); }
    px.flags_offset = at;
    at += 2ll * EX_KINDS * d.world;
    // flag-in-data slots of the two per-update exchanges: two 8-byte words per double (kernels.cuh)
    for (int parity = 0; parity < 2; parity++)
        for (int k = 0; k < 2; k++) { px.ll_offset[parity][k] = at; at += 2ll * d.world * px.count[k]; }
    CUDA_TRY(e, cudaMalloc(&e->mailbox, (size_t)at * sizeof(double)));
    CUDA_TRY(e, cudaMemset(e->mailbox, 0, (size_t)at * sizeof(double)));
    CUDA_TRY(e, cudaHostAlloc(&e->h_p2p_error, sizeof(int), cudaHostAllocMapped));
    *e->h_p2p_error = 0;
    CUDA_TRY(e, cudaHostGetDevicePointer((void **)&px.error, e->h_p2p_error, 0));
    px.copies_done = dev_alloc<int>(e, 1);
    px.error_dev = dev_alloc<int>(e, 1);
    if (!px.copies_done || !px.error_dev) return fail(e, MPPI_B200_ERR_CUDA, "device allocation failed");
    px.timeout_cycles = 4000000000ll;   // ~2 s
    for (int q = 0; q < MPPI_MAX_WORLD; q++) px.mail[q] = nullptr;
    px.mail[d.rank] = e->mailbox;
    return MPPI_B200_OK;
}

int mppi_b200_p2p_handle(mppi_b200_engine *e, void *handle64) {
    if (!e || !handle64) return MPPI_B200_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    const int rc = p2p_allocate(e);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CUDA_TRY(e, cudaIpcGetMemHandle(&h, e->mailbox));
    std::memcpy(handle64, &h, 64);
    return MPPI_B200_OK;
}

int mppi_b200_p2p_init(mppi_b200_engine *e, const void *handles) {
    if (!e || !handles) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    const int rc = p2p_allocate(e);
    if (rc) return rc;
    if (e->comm) return fail(e, MPPI_B200_ERR_INVALID, "this engine already exchanges through NCCL");
    for (int q = 0; q < e->d.world; q++) {
        if (q == e->d.rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const unsigned char *>(handles) + 64 * (size_t)q, 64);
        void *ptr = nullptr;
        CUDA_TRY(e, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        e->peer_mappings.push_back(ptr);
        e->px.mail[q] = static_cast<double *>(ptr);
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    for (cudaGraphExec_t &g : e->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    e->d.px = e->px;       // by value in the kernels' parameter bank
    e->d.has_px = 1;
    e->p2p = true;
    return MPPI_B200_OK;
}

// mppi.cpp:481-512
int mppi_b200_get(mppi_b200_engine *e, double *control, double time) {
    if (!e || !control) return MPPI_B200_ERR_INVALID;
    const int nu = e->d.nu, T = e->d.T;
    std::lock_guard<std::mutex> publish(e->publish_mutex);
    if (time < e->last_rollout_time) return MPPI_B200_ERR_INVALID;   // "time >= m_last_rollout_time" (assert, mppi.cpp:483); no message: e->error belongs to the updating thread
    const double t0 = (time - e->last_rollout_time) / e->d.dt;
    const int lower = (int)t0, upper = lower + 1;
    for (int c = 0; c < e->batch; c++) {
        const double *U = e->h_U + (size_t)c * nu * T;
        double *out = control + (size_t)c * nu;
        if (upper >= T) {
            for (int i = 0; i < nu; i++) out[i] = e->has_default ? e->control_default[i] : U[(size_t)(T - 1) * nu + i];
            continue;
        }
        const double t = t0 - lower;
        for (int i = 0; i < nu; i++) out[i] = (1.0 - t) * U[(size_t)lower * nu + i] + t * U[(size_t)upper * nu + i];
    }
    return MPPI_B200_OK;
}

int mppi_b200_read(mppi_b200_engine *e, int32_t what, void *dst, size_t bytes) {
    if (!e || !dst) return MPPI_B200_ERR_INVALID;
    DeviceState &d = e->d;
    int rc = mppi_b200_synchronize(e);
    if (rc) return rc;
    const size_t B = (size_t)e->batch, n = (size_t)d.nu * d.T, K = (size_t)d.k_count;
    auto need = [&](size_t b) { return bytes == b; };
    switch (what) {
        case MPPI_B200_READ_OPTIMAL: if (!need(B * n * 8)) break; std::memcpy(dst, e->h_U, bytes); return MPPI_B200_OK;
        case MPPI_B200_READ_COSTS: if (!need(B * K * 8)) break; CUDA_TRY(e, cudaMemcpy(dst, d.costs, bytes, cudaMemcpyDeviceToHost)); return MPPI_B200_OK;
        case MPPI_B200_READ_WEIGHTS: {
            // the device keeps exp(-cost_scale (c - min)/(max - min)); the division by the total of
            // mppi.cpp:403-408 is applied here (the weighted sum divides once, in k_finish)
            if (!need(B * K * 8)) break;
            CUDA_TRY(e, cudaMemcpy(dst, d.weights, bytes, cudaMemcpyDeviceToHost));
            double *w = static_cast<double *>(dst);
            for (size_t c = 0; c < B; c++)
                if (e->weights_valid[c]) for (size_t i = 0; i < K; i++) w[c * K + i] = w[c * K + i] / e->weights_total[c];
            return MPPI_B200_OK;
        }
        case MPPI_B200_READ_GRADIENT: if (!need(B * n * 8)) break; CUDA_TRY(e, cudaMemcpy(dst, d.gradient, bytes, cudaMemcpyDeviceToHost)); return MPPI_B200_OK;
        case MPPI_B200_READ_NOISE: {
            if (!need(B * K * n * 8)) break;
            if (e->cfg.precision == MPPI_B200_FP64) { CUDA_TRY(e, cudaMemcpy(dst, d.noise, bytes, cudaMemcpyDeviceToHost)); return MPPI_B200_OK; }
            std::vector<float> tmp(B * K * n);
            CUDA_TRY(e, cudaMemcpy(tmp.data(), d.noise, tmp.size() * 4, cudaMemcpyDeviceToHost));
            double *o = static_cast<double *>(dst);
            for (size_t i = 0; i < tmp.size(); i++) o[i] = (double)tmp[i];
            return MPPI_B200_OK;
        }
        case MPPI_B200_READ_MINMAX: {
            if (!need(B * 16)) break;
            for (size_t c = 0; c < B; c++) { static_cast<double *>(dst)[2 * c] = -e->h_stats[5 * c]; static_cast<double *>(dst)[2 * c + 1] = e->h_stats[5 * c + 1]; }
            return MPPI_B200_OK;
        }
        // the re-rollout of the LAST update, evaluated by the first read after it (run_optimal)
        case MPPI_B200_READ_OPTIMAL_COST:
        case MPPI_B200_READ_BREAKDOWN: {
            if (!need(what == MPPI_B200_READ_OPTIMAL_COST ? B * 8 : B * 64)) break;
            if (e->snapshot_valid && e->optimal_for != e->update_count) { rc = run_optimal(e); if (rc) return rc; }
            std::memcpy(dst, e->h_opt + (what == MPPI_B200_READ_OPTIMAL_COST ? 0 : B), bytes);
            return MPPI_B200_OK;
        }
        case MPPI_B200_READ_KEPT: {
            const size_t k = bytes / 8;
            if (bytes % 8 || k > B * (size_t)std::max<long long>(d.keep_best, 1)) break;
            CUDA_TRY(e, cudaMemcpy(dst, d.kept_list, bytes, cudaMemcpyDeviceToHost));
            return MPPI_B200_OK;
        }
    }
    return fail(e, MPPI_B200_ERR_INVALID, "read: unknown item or wrong size");
}

int mppi_b200_query(mppi_b200_engine *e, int32_t what, int64_t *value) {
    if (!e || !value) return MPPI_B200_ERR_INVALID;
    switch (what) {
        case MPPI_B200_QUERY_STEP_COUNT: *value = e->d.T; return 0;
        case MPPI_B200_QUERY_ROLLOUT_COUNT: *value = e->d.K_total; return 0;
        case MPPI_B200_QUERY_LOCAL_BEGIN: *value = e->d.k_begin; return 0;
        case MPPI_B200_QUERY_LOCAL_COUNT: *value = e->d.k_count; return 0;
        case MPPI_B200_QUERY_UPDATE_COUNT: *value = e->update_count; return 0;
        case MPPI_B200_QUERY_KERNEL_LAUNCHES: *value = e->launches; return 0;
        case MPPI_B200_QUERY_ARGMIN: *value = e->argmin[0]; return 0;
        case MPPI_B200_QUERY_BATCH: *value = e->batch; return 0;
        case MPPI_B200_QUERY_SHIFT_BY: *value = e->shift_by; return 0;
        case MPPI_B200_QUERY_STATE_DOF: *value = e->d.nx; return 0;
        case MPPI_B200_QUERY_CONTROL_DOF: *value = e->d.nu; return 0;
    }
    return MPPI_B200_ERR_INVALID;
}

int mppi_b200_set_profiling(mppi_b200_engine *e, int32_t enabled) {
    if (!e) return MPPI_B200_ERR_INVALID;
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    if (enabled) for (cudaEvent_t &ev : e->ev_stage) if (!ev) CUDA_TRY(e, cudaEventCreate(&ev));
    e->profiling = enabled != 0;
    return MPPI_B200_OK;
}

int mppi_b200_stage_seconds(mppi_b200_engine *e, double *seconds, size_t count) {
    if (!e || !seconds || count != MPPI_B200_STAGES) return MPPI_B200_ERR_INVALID;
    std::memcpy(seconds, e->stage_s, sizeof e->stage_s);
    return MPPI_B200_OK;
}

int mppi_b200_measure_fma_peak(int32_t device, int32_t precision, double *tflops) {
    if (!tflops) return MPPI_B200_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return MPPI_B200_ERR_CUDA;
    double v = 0.0;
    if (measure_fma_peak(precision, &v) != cudaSuccess) return MPPI_B200_ERR_CUDA;
    *tflops = v;
    return MPPI_B200_OK;
}

int mppi_b200_last_update_device_seconds(mppi_b200_engine *e, double *seconds) {
    if (!e || !seconds) return MPPI_B200_ERR_INVALID;
    if (e->timing_pending) {
        if (cudaEventSynchronize(e->ev_end) == cudaSuccess) cudaEventElapsedTime(&e->last_ms, e->ev_start, e->ev_end);
        e->timing_pending = false;
    }
    *seconds = (double)e->last_ms * 1e-3;
    return MPPI_B200_OK;
}

static mppi_b200_barrier B(double b, double s) { mppi_b200_barrier r; r.bound = b; r.scale = s; r.maximum_cost = 1e10; return r; }
static mppi_b200_quadratic Q(double c0, double c1, double c2) { mppi_b200_quadratic r; r.constant_cost = c0; r.linear_cost = c1; r.quadratic_cost = c2; return r; }
static const double ARM_LOWER[12] = {-2.0, -2.0, -6.28, -2.8, -1.745, -2.8, -3.0718, -2.7925, 0.349, -2.967, 0.0, 0.0};
static const double ARM_UPPER[12] = {2.0, 2.0, 6.28, 2.8, 1.745, 2.8, 0.0, 2.7925, 4.53785, 2.967, 0.5, 0.5};

void mppi_b200_default_toy_objective(mppi_b200_toy_objective *o) {
    o->target[0] = 1.0; o->target[1] = 1.0; o->position_cost = 100.0; o->velocity_cost = 1.0; o->control_cost = 0.01;
}

// track_point.hpp:77-114
void mppi_b200_default_track_point(mppi_b200_track_point *o) {
    std::memset(o, 0, sizeof *o);
    o->point[0] = o->point[1] = o->point[2] = 1.0;
    o->enable_joint_limits = 1;
    static const double lo_s[12] = {1.0, 0.0, 0.0, 10.0, 50.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0};
    static const double hi_s[12] = {0.0, 0.0, 0.0, 10.0, 50.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0};
    for (int i = 0; i < 12; i++) { o->lower_joint_limit[i] = B(ARM_LOWER[i], lo_s[i]); o->upper_joint_limit[i] = B(ARM_UPPER[i], hi_s[i]); }
    o->self_collision_limit = B(0.0, 1.0);
    o->self_collision_radii[0] = 0.75;
    for (int i = 1; i < 8; i++) o->self_collision_radii[i] = 0.1;
    o->maximum_reach_limit = B(0.8, 1.0);
    o->link_position_mode = MPPI_B200_LINKS_ZERO;
}

// assisted_manipulation.hpp:133-206
void mppi_b200_default_assisted_manipulation(mppi_b200_assisted_manipulation *o) {
    std::memset(o, 0, sizeof *o);
    o->enable_joint_limit = o->enable_self_collision_limit = o->enable_workspace_limit = 1;
    o->enable_energy_limit = 0;
    o->enable_velocity_cost = o->enable_trajectory_cost = o->enable_manipulability_cost = 1;
    o->link_position_mode = MPPI_B200_LINKS_ZERO;
    static const double s[12] = {0.0, 0.0, 0.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 0.0, 0.0};
    static const double vq[12] = {1000.0, 1000.0, 100.0, 0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.0, 0.0};
    for (int i = 0; i < 12; i++) { o->lower_joint_limit[i] = B(ARM_LOWER[i], s[i]); o->upper_joint_limit[i] = B(ARM_UPPER[i], s[i]); o->velocity_cost[i] = Q(0, 0, vq[i]); }
    o->self_collision_limit = B(0.0, 1.0);
    o->self_collision_radii[0] = 0.75;
    for (int i = 1; i < 8; i++) o->self_collision_radii[i] = 0.1;
    o->workspace_limit_above = B(0.0, 1.0); o->workspace_limit_infront = B(0.0, 1.0); o->workspace_limit_reach = B(1.0, 1.0);
    o->workspace_cost_yaw = Q(0, 0, 400.0);
    o->energy_limit_below = B(0.0, 10.0); o->energy_limit_above = B(20.0, 10.0);
    o->trajectory_target_scale = 1e-2; o->trajectory_target_maximum = 1.0;
    o->trajectory_position_cost = Q(100.0, 0, 500.0); o->trajectory_position_threshold = 0.0;
    o->trajectory_velocity_cost = Q(0, 0, 500.0);
    o->trajectory_velocity_minimum = 0.1; o->trajectory_velocity_maximum = 5.0; o->trajectory_velocity_dropoff = 2.0;
    o->manipulability_cost = Q(0, 0, 10.0);
}

}  // extern "C"
