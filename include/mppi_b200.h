/*
 * mppi_b200.h — C ABI of the B200-native MPPI rollout engine.
 *
 * Drop-in boundary for ONE path of LuigiVan01/AssistedManipulation: mppi::Trajectory::update()
 * (reference src/controller/mppi.cpp:154-187) and everything it calls per sample. The
 * reference has no FFI of its own (it is a single C++ executable), so these entry points are
 * what a reference-side `mppi::Trajectory` would bind instead of its private
 * sample()/rollout()/optimise()/filter() members; the C++ facade in
 * assistedmanipulation_b200/cpp/mppi_b200/trajectory.hpp does exactly that and keeps the
 * reference's public API. INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes only; every function returns MPPI_B200_OK (0) or a
 * negative error code and never throws; matrices are column-major like Eigen
 * (mppi.hpp:262-265): element (d, t) of a nu x T matrix is at d + nu * t. An engine owns all
 * of its device memory and streams; it is not thread-safe, distinct engines are independent.
 * There is no CPU fallback: if no CUDA device is usable, create fails.
 */
#ifndef MPPI_B200_H
#define MPPI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_B200_ABI_VERSION 2

/* error codes */
#define MPPI_B200_OK 0
#define MPPI_B200_ERR_INVALID (-1)       /* bad argument / configuration (mppi.cpp:18-69 conditions) */
#define MPPI_B200_ERR_UNSUPPORTED (-2)   /* unknown (dynamics, cost) pair: there is no CPU fallback */
#define MPPI_B200_ERR_CUDA (-3)          /* CUDA runtime failure, see mppi_b200_last_error */
#define MPPI_B200_ERR_ALL_NAN (-4)       /* std::runtime_error("all nan rollouts"), mppi.cpp:368-370 */
#define MPPI_B200_ERR_TIME (-5)          /* smoothing window reset into the past, filter.cpp:37-44 */
#define MPPI_B200_ERR_NCCL (-6)

/* systems: the concrete mppi::Dynamics implementations a device binding exists for */
#define MPPI_B200_SYSTEM_TOY 0               /* 2-D double integrator (BASELINE.json config 1; new) */
#define MPPI_B200_SYSTEM_FRANKA_RIDGEBACK 1  /* FrankaRidgeback::PinocchioDynamics, pinocchio_dynamics.cpp:142-260 */

/* objectives: the concrete mppi::Cost implementations */
#define MPPI_B200_OBJECTIVE_TOY 0
#define MPPI_B200_OBJECTIVE_TRACK_POINT 1            /* objective/track_point.cpp:10-174 */
#define MPPI_B200_OBJECTIVE_ASSISTED_MANIPULATION 2  /* objective/assisted_manipulation.cpp:37-319 */

/* arithmetic of the rollout kernel (costs, weights and the update are always FP64) */
#define MPPI_B200_FP64 0
#define MPPI_B200_FP32 1

/* how the rigid body step is evaluated (results agree to rounding) */
#define MPPI_B200_DYNAMICS_FAITHFUL 0  /* tau = u + nle(q,v); a = ABA(q,v,tau) as pinocchio_dynamics.cpp:156-171 */
#define MPPI_B200_DYNAMICS_FUSED 1     /* a = M(q)^-1 u (the nle terms cancel analytically); nle only where the tank needs it */

/* Pinocchio backend returns zero for every link position (pinocchio_dynamics.hpp:189-192) */
#define MPPI_B200_LINKS_ZERO 0      /* reference (Pinocchio backend) behaviour */
#define MPPI_B200_LINKS_BODY_COM 1  /* RaiSim intent: body centre of mass from FK (raisim_dynamics.hpp:223-228) */

/* where the injected noise lives */
#define MPPI_B200_NOISE_PHILOX 0  /* generate in-kernel: Philox4x32-10, key = seed, counter = (update, k, t, d/4) */
#define MPPI_B200_NOISE_HOST 1    /* inject from a host buffer  [(K+2)][T][nu] (nu fastest), FP64 */
#define MPPI_B200_NOISE_DEVICE 2  /* inject from a device buffer, same layout, engine precision */

/* cost.hpp:43-99 — Left/RightInverseBarrierFunction {bound, scale, maximum_cost} */
typedef struct mppi_b200_barrier {
    double bound;
    double scale;
    double maximum_cost;
} mppi_b200_barrier;

/* cost.hpp:10-37 — QuadraticCost */
typedef struct mppi_b200_quadratic {
    double constant_cost;
    double linear_cost;
    double quadratic_cost;
} mppi_b200_quadratic;

/* toy objective (BASELINE.json config 1): q_p |p - target|^2 + q_v |v|^2 + q_u |u|^2 */
typedef struct mppi_b200_toy_objective {
    double target[2];
    double position_cost, velocity_cost, control_cost;
} mppi_b200_toy_objective;

/* FrankaRidgeback::TrackPoint::Configuration, objective/track_point.hpp:20-60 */
typedef struct mppi_b200_track_point {
    double point[3];
    int32_t enable_joint_limits;
    int32_t enable_self_collision_avoidance;
    int32_t enable_power_limit; /* carried; no reference term reads it */
    int32_t enable_reach_limits;
    mppi_b200_barrier lower_joint_limit[12]; /* carried; track_point.cpp:45-79 hard-codes its limits */
    mppi_b200_barrier upper_joint_limit[12];
    mppi_b200_barrier self_collision_limit;
    double self_collision_radii[8];
    mppi_b200_barrier maximum_reach_limit;
    int32_t link_position_mode; /* MPPI_B200_LINKS_* */
    int32_t reserved;
} mppi_b200_track_point;

/* FrankaRidgeback::AssistedManipulation::Configuration, objective/assisted_manipulation.hpp:20-99 */
typedef struct mppi_b200_assisted_manipulation {
    int32_t enable_joint_limit;
    int32_t enable_self_collision_limit;
    int32_t enable_workspace_limit;
    int32_t enable_energy_limit;
    int32_t enable_velocity_cost;
    int32_t enable_trajectory_cost;
    int32_t enable_manipulability_cost;
    int32_t link_position_mode; /* MPPI_B200_LINKS_* */
    mppi_b200_barrier lower_joint_limit[12];
    mppi_b200_barrier upper_joint_limit[12];
    mppi_b200_barrier self_collision_limit;
    double self_collision_radii[8];
    mppi_b200_barrier workspace_limit_above;
    mppi_b200_barrier workspace_limit_infront;
    mppi_b200_barrier workspace_limit_reach;
    mppi_b200_quadratic workspace_cost_yaw;
    mppi_b200_barrier energy_limit_below;
    mppi_b200_barrier energy_limit_above;
    mppi_b200_quadratic velocity_cost[12];
    double trajectory_target_scale;
    double trajectory_target_maximum;
    mppi_b200_quadratic trajectory_position_cost;
    double trajectory_position_threshold;
    mppi_b200_quadratic trajectory_velocity_cost;
    double trajectory_velocity_minimum;
    double trajectory_velocity_maximum;
    double trajectory_velocity_dropoff;
    mppi_b200_quadratic manipulability_cost;
} mppi_b200_assisted_manipulation;

/* mppi::Configuration (mppi.hpp:181-249) + what the device binding needs to know */
typedef struct mppi_b200_config {
    int32_t abi_version; /* MPPI_B200_ABI_VERSION */
    int32_t system;      /* MPPI_B200_SYSTEM_* */
    int32_t objective;   /* MPPI_B200_OBJECTIVE_* */
    int32_t precision;   /* MPPI_B200_FP64 / FP32 */
    int32_t dynamics_mode; /* MPPI_B200_DYNAMICS_* */
    int32_t device;      /* CUDA ordinal */
    int32_t rank;        /* this engine owns global samples [rank*(K+2)/world, (rank+1)*(K+2)/world) */
    int32_t world_size;
    int32_t state_dof;   /* dynamics->get_state_dof(), checked against the system */
    int32_t control_dof; /* dynamics->get_control_dof() */
    int64_t rollouts;    /* K; the engine runs K + 2 (mppi.hpp:306) */
    int64_t keep_best_rollouts;
    double time_step;
    double horison;      /* sic, mppi.hpp:196 */
    double gradient_step;
    double cost_scale;
    double cost_discount_factor;
    const double *covariance; /* nu x nu, column-major */
    int32_t covariance_rows, covariance_cols;
    int32_t control_bound;
    int32_t control_limits_size; /* length of control_min / control_max */
    const double *control_min;
    const double *control_max;
    const double *control_default; /* nu values or NULL (std::nullopt) */
    int32_t smoothing;             /* 0 = std::nullopt */
    uint32_t smoothing_window;
    uint32_t smoothing_order;
    int32_t threads; /* validated like mppi.cpp:66 (> 0); the CUDA grid replaces the pool */
    /* Number of independent controllers that share this configuration and run as ONE grid (BASELINE.json
     * config 5). 0 or 1 = a single mppi::Trajectory. With batch = B every per-update input and every
     * read-back is the concatenation over controllers: state B x state_dof, wrench B x T x 6, noise
     * B x (K+2) x T x nu, mppi_b200_get B x control_dof, ...; controller c draws Philox noise with seed + c. */
    int32_t batch;
} mppi_b200_config;

typedef struct mppi_b200_engine mppi_b200_engine;

/* mppi::Trajectory::create, mppi.cpp:11-77. objective_params points at the struct matching
 * config->objective. On failure *engine is NULL and mppi_b200_last_error(NULL) has the reason. */
int mppi_b200_create(const mppi_b200_config *config, const void *objective_params, size_t objective_params_size,
                     mppi_b200_engine **engine);
void mppi_b200_destroy(mppi_b200_engine *engine);
const char *mppi_b200_last_error(const mppi_b200_engine *engine);

/* mppi::Trajectory::update, mppi.cpp:154-187: sample -> rollout -> optimise -> (re-rollout).
 *   state  : host, state_dof doubles.
 *   wrench : host, T x 6 doubles — forecast->get_end_effector_wrench(time + t*dt).head(6) for each
 *            horizon step (dynamics.hpp:275-278); NULL = no forecast handle (trajectory term = 0).
 *   noise  : per noise_source; host buffers are copied inside the call.
 * Returns after the updated control sequence is resident on the host (mppi_b200_get is valid). */
int mppi_b200_update(mppi_b200_engine *engine, const double *state, double time, const double *wrench,
                     const void *noise, int32_t noise_source, uint64_t seed);

/* Asynchronous form of mppi_b200_update for many independent controllers on one device (BASELINE.json
 * config 5): _launch enqueues the whole update on the engine's own stream and returns; _wait blocks until
 * that engine's control sequence is on the host. Launch all engines, then wait for all: their updates
 * overlap on the GPU. Exactly one _wait per _launch. */
int mppi_b200_update_launch(mppi_b200_engine *engine, const double *state, double time, const double *wrench,
                            const void *noise, int32_t noise_source, uint64_t seed);
int mppi_b200_update_wait(mppi_b200_engine *engine);

/* The same update split at its two exchange points, for rollout sets sharded over engines
 * (one per GPU): begin -> [all-reduce MAX of minmax buffer] -> weights -> [all-reduce SUM of sums
 * buffer] -> finish. All asynchronous on the engine stream until finish returns. */
int mppi_b200_update_begin(mppi_b200_engine *engine, const double *state, double time, const double *wrench,
                           const void *noise, int32_t noise_source, uint64_t seed);
int mppi_b200_update_weights(mppi_b200_engine *engine);
int mppi_b200_update_finish(mppi_b200_engine *engine);
/* device addresses of the two exchange buffers (FP64). minmax = {-min, max, this rank's valid count, and when sharded one slot per rank: that
 * rank's count of valid (non-NaN) rollouts, saturated at 2, others 0} — all-reduce all `minmax_count` values with MAX (the
 * slots are disjoint, so MAX gathers them; the engine sums them: two ranks with one valid rollout each are two valid
 * rollouts, mppi.cpp:368-370). sums = {sum w, sum w*eps[nu*T], and when sharded one slot per rank: that rank's best global
 * rollout index + 1, or 0} — all-reduce all `sums_count` values with SUM. */
int mppi_b200_reduce_buffers(mppi_b200_engine *engine, void **minmax, size_t *minmax_count, void **sums,
                             size_t *sums_count);
int mppi_b200_stream(mppi_b200_engine *engine, void **cuda_stream);
int mppi_b200_synchronize(mppi_b200_engine *engine);

/* Optional in-library exchange over NCCL (NVLink 5 / NVSwitch): with a communicator attached,
 * mppi_b200_update runs the two all-reduces itself on the engine stream. */
int mppi_b200_comm_unique_id(void *id128);
int mppi_b200_comm_init(mppi_b200_engine *engine, const void *id128);

/* In-library exchange over NVLink peer memory, one process per GPU (preferred to NCCL for this path: the payloads are
 * 24 B and ~6 KB, the cost is latency). Every rank exports its mailbox (a cudaIpcMemHandle_t, 64 bytes); the caller
 * gathers the handles of all ranks in rank order with whatever transport it has and hands them to every engine. After
 * that mppi_b200_update performs the min/max, weighted-sum and warm-start exchanges with its own kernels (stores into
 * the peers' mailboxes, flags, rank-ordered combine), all inside the update's CUDA graph. A rank that does not arrive
 * within ~2 s makes the update return MPPI_B200_ERR_NCCL instead of hanging the device. */
int mppi_b200_p2p_handle(mppi_b200_engine *engine, void *handle64);
int mppi_b200_p2p_init(mppi_b200_engine *engine, const void *handles /* world_size x 64 bytes, rank order */);

/* mppi::Trajectory::get, mppi.cpp:481-512 (host side, linear interpolation / default control).
 * Thread safety: the one call that may run on ANOTHER thread while mppi_b200_update / _update_launch / _update_wait /
 * _update_finish run on the engine's owner thread (the reference's control loop reads while the controller updates,
 * mppi.cpp:179,492). It sees either the previously published sequence or the new one, never a mixture: the engine
 * serialises it against the publication step only (the copy of the sequence and of its time stamp), not against the
 * update. A time before the last published update returns MPPI_B200_ERR_INVALID (the reference asserts) without touching
 * the engine's error message. Every other entry point is single-threaded per engine; distinct engines are independent. */
int mppi_b200_get(mppi_b200_engine *engine, double *control, double time);

/* read-back of the getters logger::MPPI and BaseTest use (logging/mppi.cpp:84-136) */
#define MPPI_B200_READ_OPTIMAL 0       /* nu x T doubles: trajectory() / get_optimal_rollout() */
#define MPPI_B200_READ_COSTS 1         /* local rollouts doubles: get_rollouts()[k].cost */
#define MPPI_B200_READ_WEIGHTS 2       /* local rollouts doubles: get_weights() */
#define MPPI_B200_READ_GRADIENT 3      /* nu x T doubles: get_gradient() */
#define MPPI_B200_READ_NOISE 4         /* local rollouts x T x nu doubles (widened in FP32 mode): get_rollouts()[k].noise */
#define MPPI_B200_READ_MINMAX 5        /* {min, max} */
#define MPPI_B200_READ_OPTIMAL_COST 6  /* get_optimal_total_cost() */
#define MPPI_B200_READ_BREAKDOWN 7     /* 8 doubles: joint, self collision, workspace, energy, velocity, trajectory, manipulability, total */
#define MPPI_B200_READ_KEPT 8          /* keep_best int64: global indices kept for the warm start, sorted order */
int mppi_b200_read(mppi_b200_engine *engine, int32_t what, void *dst, size_t bytes);

#define MPPI_B200_QUERY_STEP_COUNT 0
#define MPPI_B200_QUERY_ROLLOUT_COUNT 1  /* K + 2 */
#define MPPI_B200_QUERY_LOCAL_BEGIN 2
#define MPPI_B200_QUERY_LOCAL_COUNT 3
#define MPPI_B200_QUERY_UPDATE_COUNT 4
#define MPPI_B200_QUERY_KERNEL_LAUNCHES 5  /* kernels this engine has launched so far */
#define MPPI_B200_QUERY_ARGMIN 6           /* global index of the lowest-cost rollout (lowest index on ties) */
#define MPPI_B200_QUERY_SHIFT_BY 7
#define MPPI_B200_QUERY_STATE_DOF 8
#define MPPI_B200_QUERY_CONTROL_DOF 9
#define MPPI_B200_QUERY_BATCH 10
int mppi_b200_query(mppi_b200_engine *engine, int32_t what, int64_t *value);

/* time of the kernels of the last update on the device, seconds (CUDA events on the engine stream). mppi_b200_update returns
 * when the update's results have landed in host memory, which is a few microseconds before the stream's end-of-update event:
 * this call waits for that event first. */
int mppi_b200_last_update_device_seconds(mppi_b200_engine *engine, double *seconds);

/* Per-stage device time of the last update (CUDA events between the stages; off by default because
 * the extra events cost a little latency). Stages: 0 host->device inputs, 1 warm start + shift,
 * 2 sample (K1), 3 rollout (K2), 4 min/max + weights (K3), 5 weighted sum (K4), 6 finish (K5),
 * 7 device->host result. */
#define MPPI_B200_STAGES 8
int mppi_b200_set_profiling(mppi_b200_engine *engine, int32_t enabled);
int mppi_b200_stage_seconds(mppi_b200_engine *engine, double *seconds, size_t count);

/* Roofline denominator for the rollout kernel: sustained FMA rate of the vector pipe of `device` in
 * the given precision (MPPI_B200_FP64 / FP32), measured with a register-resident FMA-chain kernel. */
int mppi_b200_measure_fma_peak(int32_t device, int32_t precision, double *tflops);

/* Defaults of the reference, so that callers need not restate them:
 * TrackPoint::DEFAULT_CONFIGURATION (track_point.hpp:77-114),
 * AssistedManipulation::DEFAULT_CONFIGURATION (assisted_manipulation.hpp:133-206). */
void mppi_b200_default_track_point(mppi_b200_track_point *out);
void mppi_b200_default_assisted_manipulation(mppi_b200_assisted_manipulation *out);
void mppi_b200_default_toy_objective(mppi_b200_toy_objective *out);

/* ---- Wrench forecast producer: the table W[t] the objective consumes --------------------
 * Batched device implementation of the reference's forecasters for 6-component wrenches
 * (src/controller/forecast.{hpp,cpp}: LOCFForecast :62-140, AverageForecast cpp:41-128, KalmanForecast
 * cpp:130-367 over KalmanFilter kalman.cpp:89-152). One object holds `batch` independent forecasters that
 * share a configuration; every call acts on all of them. */
#define MPPI_B200_FORECAST_LOCF 0     /* Forecast::Configuration::Type, forecast.hpp:391-396 */
#define MPPI_B200_FORECAST_AVERAGE 1
#define MPPI_B200_FORECAST_KALMAN 2

typedef struct mppi_b200_forecast_config {
    int32_t type;      /* MPPI_B200_FORECAST_* */
    int32_t batch;     /* forecasters (>= 1) */
    int32_t device;
    uint32_t order;    /* KalmanForecast::Configuration::order (0..2) */
    double time_step;  /* KalmanForecast::Configuration::time_step */
    double horison;    /* LOCF / Kalman horison */
    double window;     /* AverageForecast::Configuration::window */
} mppi_b200_forecast_config;

typedef struct mppi_b200_forecast mppi_b200_forecast;

/* initial: batch x 6 (LOCF observation / Kalman initial_state) or NULL for zeros */
int mppi_b200_forecast_create(const mppi_b200_forecast_config *config, const double *initial, mppi_b200_forecast **forecast);
void mppi_b200_forecast_destroy(mppi_b200_forecast *forecast);
/* message of the last failed call on this object (NULL: of the last failed create on this thread) */
const char *mppi_b200_forecast_last_error(const mppi_b200_forecast *forecast);
/* Forecast::update(measurement, time) for every forecaster; measurements: host, batch x 6 */
int mppi_b200_forecast_update(mppi_b200_forecast *forecast, const double *measurements, double time);
/* Forecast::update(time) */
int mppi_b200_forecast_update_time(mppi_b200_forecast *forecast, double time);
/* Forecast::forecast(time + k * time_step).head(6) for k < steps: host table, batch x steps x 6 */
int mppi_b200_forecast_table(mppi_b200_forecast *forecast, double time, double time_step, int32_t steps, double *table);
/* The same table left on the device (valid until the next call on this object); hand it to
 * mppi_b200_set_wrench_device so a batched engine consumes it without a host round trip. */
int mppi_b200_forecast_table_device(mppi_b200_forecast *forecast, double time, double time_step, int32_t steps, const double **device_table);
/* From now on the engine takes its forecast wrench table (batch x T x 6, FP64) from this device address instead
 * of the `wrench` argument of mppi_b200_update; NULL restores the host argument. */
int mppi_b200_set_wrench_device(mppi_b200_engine *engine, const double *device_table);

/* forecasters in the object */
int mppi_b200_forecast_batch(const mppi_b200_forecast *forecast);

/* ---- Dynamics forecast: FrankaRidgeback::DynamicsForecast::forecast ---------------------------
 * (src/frankaridgeback/dynamics.{hpp:122-408,cpp:58-138}) for `batch` controllers at once: the Pinocchio-backend
 * dynamics rolled forward under zero control over ceil(horison / time_step) steps, one record per step taken
 * BEFORE the step. Record layout (doubles): joint_position[12], end effector position[3], orientation
 * quaternion x y z w [4], linear_velocity[3], angular_velocity[3], linear_acceleration[3],
 * angular_acceleration[3], joint_power, external_power, energy, wrench[6], jacobian[6][12] row-major. */
#define MPPI_B200_DYNAMICS_FORECAST_RECORD 112

typedef struct mppi_b200_dynamics_forecast_config {
    int32_t batch;
    int32_t device;
    double time_step;      /* DynamicsForecast::Configuration::time_step */
    double horison;        /* DynamicsForecast::Configuration::horison */
    int32_t apply_wrench;  /* 0 = the reference (add_end_effector_simulated_wrench is empty in the Pinocchio backend,
                              pinocchio_dynamics.hpp:276); 1 = tau += J_ee^T w, the line left commented out at
                              pinocchio_dynamics.cpp:240, fed with the forecast wrench */
    int32_t reserved;
} mppi_b200_dynamics_forecast_config;

typedef struct mppi_b200_dynamics_forecast mppi_b200_dynamics_forecast;

/* wrench_forecast: the end_effector_wrench_forecast (same batch; not owned; NULL = zero wrench) */
int mppi_b200_dynamics_forecast_create(const mppi_b200_dynamics_forecast_config *config, mppi_b200_forecast *wrench_forecast,
                                       mppi_b200_dynamics_forecast **forecast);
void mppi_b200_dynamics_forecast_destroy(mppi_b200_dynamics_forecast *forecast);
const char *mppi_b200_dynamics_forecast_last_error(const mppi_b200_dynamics_forecast *forecast);
int mppi_b200_dynamics_forecast_steps(const mppi_b200_dynamics_forecast *forecast);
/* DynamicsForecast::forecast(state, time); states: host, batch x 31 */
int mppi_b200_dynamics_forecast_run(mppi_b200_dynamics_forecast *forecast, const double *states, double time);
/* records: host, batch x steps x MPPI_B200_DYNAMICS_FORECAST_RECORD doubles */
int mppi_b200_dynamics_forecast_read(mppi_b200_dynamics_forecast *forecast, double *records, size_t bytes);
int mppi_b200_dynamics_forecast_device_records(mppi_b200_dynamics_forecast *forecast, const double **records);

/* Host-side Philox4x32-10 + Box–Muller exactly as the sampling kernel evaluates it is NOT
 * provided: the engine's generated noise is read back with MPPI_B200_READ_NOISE instead. */

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H */
