// Launch interface between the host engine (engine.cu) and the sm_100a kernels.
// Kernel list: K1 sample, K2 rollout, K3 reduce/weights, K4 weighted sum,
// K5 finish (gradient step + Savitzky–Golay smoothing + clamp), K6 optimal re-rollout (= K2 with
// one thread on the zero-noise row).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rollout_core.cuh"

namespace mppi_b200 {

constexpr int MAX_NU = 12;
constexpr int MAX_WINDOW = 64;  // Savitzky–Golay half window supported by the finish kernel
constexpr int MPPI_MAX_WORLD = 16;
constexpr int MPPI_FUSED_ROWS = 512;   // rollouts per block of the weighted-sum kernel up to which it computes their weights itself (shared memory)

// Per-update inputs, written by the host into pinned memory and copied to the device in ONE
// transfer; kernels read it from global memory so a captured CUDA graph stays valid.
struct Frame {
    double x0[32];       // state (31 used for Franka+Ridgeback, 4 for the toy)
    double time;         // m_rollout_time
    double sg_prev_trim; // last smoothing reset time before this update
    long long shift_by;  // (int64)((time - last_shift_time) / dt), evaluated on the host in double (mppi.cpp:194)
    unsigned long long seed;
    unsigned long long update_index;
    unsigned long long attempt;   // every update call, also one that ends in an error: sequence number of the peer exchange
    int has_wrench;
    int noise_source;    // MPPI_B200_NOISE_*
    // followed by T x 6 doubles of forecast wrench
};

// description of the peer-memory exchange of a sharded rollout set (see the comment further down)
enum ExchangeKind { EX_MINMAX = 0, EX_SUMS = 1, EX_CAND = 2, EX_KINDS = 3 };
struct PeerExchange {
    int world, rank;
    double *mail[MPPI_MAX_WORLD];      // mailbox of every rank as mapped in this process (mail[rank] is local)
    long long offset[2][EX_KINDS];     // doubles: start of the [world][count] slot array of (parity, kind)
    long long flags_offset;            // doubles: start of the flags, unsigned long long [2][EX_KINDS][world]
    long long ll_offset[2][2];         // doubles: start of the flag-in-data slot array of (parity, kind) for EX_MINMAX / EX_SUMS: [world][count][2] 8-byte words
    int count[EX_KINDS];               // doubles per slot
    int *error;                        // host-mapped: a peer did not arrive in time
    int *error_dev;                    // the same flag in device memory (what the kernels test: a host-mapped read is a PCIe round trip)
    int *copies_done;                  // device: copy blocks of a running k_exchange launch that have read the payload
    long long timeout_cycles;
};

// Device-resident state of one engine (pointers only; owned by the engine).
struct DeviceState {
    // geometry
    int batch;             // independent controllers sharing this engine (blockIdx.y); 1 = a single mppi::Trajectory
    int elem_bytes;        // sizeof(kernel real)
    int nu, nx, T;
    long long K_total;     // K + 2
    long long k_begin;     // first global rollout index owned by this engine
    long long k_count;     // rollouts owned
    long long keep_best;
    // buffers
    const Frame *frame;
    const double *wrench;  // inside the frame block
    double *U;             // published optimal control, nu x T
    double *U_shift;       // working copy (m_optimal_control_shifted)
    void *noise;           // k_count x T x nu, engine precision
    const void *injected;  // same layout (precision of `injected_is_double`), or nullptr
    int injected_is_double;
    double *costs;         // k_count
    double *weights;       // k_count (unnormalised until finish)
    unsigned char *kept;   // k_count flags for the warm start
    long long *kept_list;  // keep_best global indices, sorted order
    int world;             // ranks sharing the rollout set
    int rank;              // this rank (its slot in the argmin part of the sums exchange buffer)
    double *cand;          // this rank's keep_best best (cost, global index bits) pairs — all-gathered when sharded
    double *cand_all;      // world x keep_best pairs
    unsigned long long *minmax_enc;  // [2] order-preserving encodings for atomicMin / atomicMax
    int *valid_count;      // number of non-NaN local rollouts
    long long *argmin;     // global index of the best rollout
    int *finish_count;     // channel blocks of k_finish that are done (the last one publishes)
    double *minmax;        // {-min, max, valid(<=2)} of the WHOLE rollout set (written by k_weights; read by k_finish and the host)
    double *minmax_local;  // exchange buffer of a sharded set: {-min, max, own valid count, valid_0 .. valid_{world-1}} with only this rank's valid slot
                           // filled (<= 2) — combined with MAX elementwise, which gathers the slots; their sum is the valid count
    double *sums;          // exchange buffer {sum w, sum w*eps [nu*T], argmin slot per rank}: this rank's part (p2p) / all-reduced in place (NCCL, split ABI)
    int has_px;            // the peer-memory exchange is attached
    PeerExchange px;       // by value: the kernels read it from their parameter bank, not through a pointer chase
    int *rollout_done;     // blocks of the rollout grid that are done (the last one publishes / pushes the min-max payload)
    int *reduce_done;      // blocks of k_gradient_reduce that are done (the last one pushes the sums payload)
    double *wsum_partial;  // per block of the weights kernel (fused tail: per block of the weighted-sum kernel), wsum_stride per controller
    double *grad_partial;  // [grad_blocks][nu*T]; fused tail: every row channel-major, [nu][T]
    int grad_blocks;
    int weight_blocks;
    int wsum_stride;       // max(weight_blocks, grad_blocks)
    int fused_tail;        // one rank, <= MPPI_FUSED_ROWS rollouts per weighted-sum block: the weights are computed by the weighted-sum kernel's
                           // blocks (each for its own rollouts) and the partial sums are combined by k_finish - two kernels less per update
    // noise chase (rollout_core.cuh / sample_core.cuh): the blocks of a small rollout grid draw their own noise
    int chase;             // 0 = the sampling kernel runs on its own
    unsigned *chase_prepared;   // the update number (low word) of the last prepare block run by the rollout grid's first block
    double *gradient;      // normalised gradient (get_gradient())
    int *skip;             // 1 when max-min < 1e-6 (mppi.cpp:373-375): weights/gradient/U left untouched
    double *L;             // nu x nu column-major noise transform V*sqrt(Lambda) (gaussian.hpp:48-55)
    int L_is_diagonal;     // diagonal covariance (every reference configuration, base.hpp:79-83): eps_i = Ldiag[i] * z_i, Ldiag = sqrt(Sigma_ii) (host_math.h)
    double Ldiag[MAX_NU];
    // smoothing window state, per channel: uu[Lw], tt[Lw], then start_idx/last_trim in sg_meta
    double *sg_uu, *sg_tt;
    double *sg_weights;    // 2*window+1
    int sg_enabled, sg_window, sg_len;
    int *sg_started;       // per channel: 0 until the first smoothing pass (window start_idx = w), then start_idx = w + T
    // constants
    double dt, gradient_step, cost_scale, discount;
    int bound;
    double cmin[MAX_NU], cmax[MAX_NU];
    // optimal re-rollout outputs
    // snapshot of this update's inputs for the optimal re-rollout on a side stream (one of the engine's slots)
    double *frame_snap;    // copy of the frame block, written by the prepare block of k_sample
    int frame_doubles;
    double *U_snap;        // copy of the updated m_optimal_control_shifted, written by k_finish
    double *result;        // host-mapped: U [nu*T], {-min, max, valid} [3], argmin [1], sum w [1] — written by k_finish
    double *optimal_cost;  // [1]
    double *breakdown;     // [8]
};

#if defined(__CUDACC__)
// The buffers of controller c of a batched engine: every per-controller buffer is laid out
// [controller][...], so a kernel block with blockIdx.y = c works on a shifted view and is otherwise
// unaware of the batch.
__device__ __forceinline__ DeviceState controller_view(const DeviceState &g, int c) {
    if (g.batch <= 1) return g;
    DeviceState d = g;
    const size_t n = (size_t)g.nu * g.T, K = (size_t)g.k_count, keep = g.keep_best > 0 ? (size_t)g.keep_best : 1;
    d.frame = reinterpret_cast<const Frame *>(reinterpret_cast<const double *>(g.frame) + (size_t)c * g.frame_doubles);
    d.wrench = g.wrench + (size_t)c * g.frame_doubles;
    d.U = g.U + c * n; d.U_shift = g.U_shift + c * n; d.gradient = g.gradient + c * n; d.U_snap = g.U_snap + c * n;
    d.noise = static_cast<unsigned char *>(g.noise) + (size_t)c * K * n * g.elem_bytes;
    if (g.injected) d.injected = static_cast<const unsigned char *>(g.injected) + (size_t)c * K * n * (g.injected_is_double ? 8 : g.elem_bytes);
    d.costs = g.costs + c * K; d.weights = g.weights + c * K; d.kept = g.kept + c * K;
    d.kept_list = g.kept_list + c * keep;
    d.minmax_enc = g.minmax_enc + 2 * c; d.valid_count = g.valid_count + c; d.argmin = g.argmin + c; d.finish_count = g.finish_count + c;
    d.minmax = g.minmax + 4 * c; d.minmax_local = g.minmax_local + (size_t)c * (3 + MPPI_MAX_WORLD); d.sums = g.sums + c * (1 + n);
    d.rollout_done = g.rollout_done + c; d.reduce_done = g.reduce_done + c;
    d.wsum_partial = g.wsum_partial + (size_t)c * g.wsum_stride; d.grad_partial = g.grad_partial + (size_t)c * g.grad_blocks * n;
    d.skip = g.skip + c;
    d.sg_uu = g.sg_uu + (size_t)c * g.nu * g.sg_len; d.sg_tt = g.sg_tt + (size_t)c * g.nu * g.sg_len; d.sg_started = g.sg_started + c * g.nu;
    d.frame_snap = g.frame_snap + (size_t)c * g.frame_doubles;
    d.result = g.result + c * (n + 8);
    d.optimal_cost = g.optimal_cost + c; d.breakdown = g.breakdown + 8 * c;
    return d;
}
#endif

// ---- exchange between the ranks of a sharded rollout set over NVLink peer memory ---------------------------
// Every rank owns a mailbox (device memory, IPC-mapped into its peers) with, per update parity and per kind, one slot per
// rank. The two per-update exchanges ride the kernels that produce and consume their payloads — no launch of their own:
// the LAST block of the rollout grid to finish stores this rank's {-min, max, valid slots} into its slot of every peer's
// mailbox (exchange_push), and the threads of k_weights poll the elements they need in the local mailbox and combine them in
// rank order (exchange_peer). Likewise the last block of k_gradient_reduce pushes {sum w, sum w*eps, argmin slots} and the
// threads of k_finish combine them — identical bits on every rank. Transport: flag-in-data (below). Replaces two NCCL
// all-reduces (~20 us each at 40 B / 6 KB) and, against round 1, two exchange launches per update. Only the warm-start
// candidates (keep_best > 0 on a sharded set) still use the stand-alone k_exchange with its payload + fence + flag protocol.
#if defined(__CUDACC__)
// sequence number of an exchange: never 0, unique per (update attempt, kind)
__device__ __forceinline__ unsigned long long exchange_seq(unsigned long long attempt, int kind) { return attempt * 4ull + (unsigned long long)kind + 1ull; }

// ---- flag-in-data transport of the two per-update exchanges (EX_MINMAX, EX_SUMS) -------------------------------------
// Every double travels as two 8-byte words {32 payload bits | 32-bit sequence number}: an 8-byte store is atomic, so a
// word that carries this exchange's sequence number carries its payload — no fence, no separate flag, no round trip: the
// producer's stores go out and the consumers poll the very words they need (the protocol NCCL calls LL). With a payload
// store + system fence + flag store the push alone waited one NVLink round trip and the flag one more hop (measured at
// 2 GPUs: ~12 us per exchange on the critical path).
__device__ __forceinline__ unsigned long long *exchange_ll_slot(const PeerExchange &px, int owner, int kind, unsigned long long attempt, int src_rank, int e) {
    return reinterpret_cast<unsigned long long *>(px.mail[owner] + px.ll_offset[(int)(attempt & 1ull)][kind]) + ((long long)src_rank * px.count[kind] + e) * 2;
}

// ONE block (all of its threads may call; at most 256 work): this rank's payload into its slot of every peer's mailbox.
// The payload must be complete and visible to the block (the callers are "last block" epilogues behind a fence + counter
// + block barrier).
__device__ __forceinline__ void exchange_push(const PeerExchange &px, int kind, unsigned long long attempt, const double *payload) {
    const int count = px.count[kind];
    const unsigned long long seq = (exchange_seq(attempt, kind) & 0xffffffffull) << 32;
    const int workers = blockDim.x < 256 ? (int)blockDim.x : 256;
    if ((int)threadIdx.x >= workers) return;
    for (int i0 = threadIdx.x; i0 < count; i0 += 4 * workers) {
        unsigned long long bits[4];
#pragma unroll
        for (int j = 0; j < 4; j++) bits[j] = (i0 + j * workers < count) ? (unsigned long long)__double_as_longlong(__ldcg(payload + i0 + j * workers)) : 0ull;   // loads first: one round trip
        for (int p = 0; p < px.world; p++) {
            if (p == px.rank) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = i0 + j * workers;
                if (i >= count) continue;
                volatile unsigned long long *dst = exchange_ll_slot(px, p, kind, attempt, px.rank, i);
                dst[0] = (bits[j] & 0xffffffffull) | seq;
                dst[1] = (bits[j] >> 32) | seq;
            }
        }
    }
}

// element e of rank q's payload (q != rank): poll the two words in the LOCAL mailbox until both carry this exchange's
// sequence number (2 s timeout -> *px.error instead of a hung device; once a peer is known missing nothing waits again)
__device__ __forceinline__ double exchange_peer(const PeerExchange &px, int kind, unsigned long long attempt, int q, int e) {
    const volatile unsigned long long *src = exchange_ll_slot(px, px.rank, kind, attempt, q, e);
    const unsigned long long seq = exchange_seq(attempt, kind) & 0xffffffffull;
    const long long t0 = clock64();
    unsigned long long w0, w1;
    for (;;) {
        w0 = src[0]; w1 = src[1];
        if ((w0 >> 32) == seq && (w1 >> 32) == seq) break;
        if (*px.error_dev || clock64() - t0 > px.timeout_cycles) { *px.error = 1; *px.error_dev = 1; break; }
    }
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

#endif

cudaError_t launch_exchange(const DeviceState &d, const PeerExchange &px, int kind, cudaStream_t s);

// ---- programmatic dependent launch (sm_90+) ------------------------------------------------------------------------
// The kernels of an update form a chain. Launched with the programmatic-serialisation attribute, a kernel's blocks may be
// brought onto the SMs while its predecessor is still running; they wait in pdl_wait() — the first statement of every
// kernel — until the predecessor has COMPLETED and its writes are visible, so nothing about the data flow changes. A
// kernel launched without the attribute passes pdl_wait() at once.
// Measured (profiles/r2_pdl_ab.log): between the small kernels the attribute changes nothing inside a CUDA graph (199 us
// either way at config 2) - all of the gain is the rollout grid moving in under the tail of the sampling kernel
// (config 3: 1090 -> 1077 us, config 5: 872 -> 852). But blocks placed while the sampling grid still holds part of every
// SM land two or three to an SM with other SMs left empty, and for a rollout grid smaller than the machine (config 2:
// 129 blocks on 148 SMs) that costs 57 % of the kernel (291 us per update instead of 199), because its step body streams
// from L2 through one instruction-fetch path per SM. So the rollout gets the attribute only when its grid is at least
// two blocks per SM (rollout_overlap_level); MPPI_B200_PDL=2 forces it, =0 turns every attribute off.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();   // this grid's successor may start moving in
#endif
}
int pdl_level();   // MPPI_B200_PDL: 0 off, 1 (default) as described above, 2 every kernel
int rollout_overlap_level(long long blocks);   // the level from which a rollout grid of this many blocks gets the attribute
template <class... KArgs, class... Args>
inline cudaError_t launch_level(int level, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_level() >= level ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <class... KArgs, class... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    return launch_level(1, kern, grid, block, smem, s, static_cast<Args &&>(args)...);
}

#endif

cudaError_t upload_robot_model();  // once per device

cudaError_t launch_select_kept(const DeviceState &d, cudaStream_t s);
cudaError_t launch_merge_kept(const DeviceState &d, cudaStream_t s);
cudaError_t launch_sample(const DeviceState &d, int precision, cudaStream_t s, int *launches);   // (d.chase: the warm-start shift only)
// objective params: pointer to the host-side block in kernel arithmetic (ToyP/TrackPointP/AssistedP<R>)
cudaError_t launch_rollout(const DeviceState &d, int precision, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query = nullptr);
cudaError_t launch_weights(const DeviceState &d, cudaStream_t s);
cudaError_t launch_gradient(const DeviceState &d, int precision, cudaStream_t s, int *launches);
cudaError_t launch_finish(const DeviceState &d, cudaStream_t s);
cudaError_t measure_fma_peak(int precision, double *tflops);

}  // namespace mppi_b200
