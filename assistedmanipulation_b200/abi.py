"""ctypes mirror of include/mppi_b200.h — the C ABI of the B200 MPPI rollout engine.

Used by tests/ and bench.py to drive the shared library through exactly the entry points a
reference-side `mppi::Trajectory` would bind (reference src/controller/mppi.cpp:154-187).
Nothing here computes: it declares structs, prototypes and the reference's default
configurations (src/test/case/base.hpp:68-100).
"""
import ctypes as C
import math
import os

import numpy as np

ABI_VERSION = 2

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_ALL_NAN, ERR_TIME, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
SYSTEM_TOY, SYSTEM_FRANKA_RIDGEBACK = 0, 1
OBJECTIVE_TOY, OBJECTIVE_TRACK_POINT, OBJECTIVE_ASSISTED_MANIPULATION = 0, 1, 2
FP64, FP32 = 0, 1
DYNAMICS_FAITHFUL, DYNAMICS_FUSED = 0, 1
LINKS_ZERO, LINKS_BODY_COM = 0, 1
NOISE_PHILOX, NOISE_HOST, NOISE_DEVICE = 0, 1, 2
(READ_OPTIMAL, READ_COSTS, READ_WEIGHTS, READ_GRADIENT, READ_NOISE, READ_MINMAX, READ_OPTIMAL_COST,
 READ_BREAKDOWN, READ_KEPT) = range(9)
(QUERY_STEP_COUNT, QUERY_ROLLOUT_COUNT, QUERY_LOCAL_BEGIN, QUERY_LOCAL_COUNT, QUERY_UPDATE_COUNT,
 QUERY_KERNEL_LAUNCHES, QUERY_ARGMIN, QUERY_SHIFT_BY, QUERY_STATE_DOF, QUERY_CONTROL_DOF, QUERY_BATCH) = range(11)


class Barrier(C.Structure):
    _fields_ = [("bound", C.c_double), ("scale", C.c_double), ("maximum_cost", C.c_double)]


class Quadratic(C.Structure):
    _fields_ = [("constant_cost", C.c_double), ("linear_cost", C.c_double), ("quadratic_cost", C.c_double)]


class ToyObjective(C.Structure):
    _fields_ = [("target", C.c_double * 2), ("position_cost", C.c_double), ("velocity_cost", C.c_double),
                ("control_cost", C.c_double)]


class TrackPoint(C.Structure):
    _fields_ = [("point", C.c_double * 3),
                ("enable_joint_limits", C.c_int32), ("enable_self_collision_avoidance", C.c_int32),
                ("enable_power_limit", C.c_int32), ("enable_reach_limits", C.c_int32),
                ("lower_joint_limit", Barrier * 12), ("upper_joint_limit", Barrier * 12),
                ("self_collision_limit", Barrier), ("self_collision_radii", C.c_double * 8),
                ("maximum_reach_limit", Barrier),
                ("link_position_mode", C.c_int32), ("reserved", C.c_int32)]


class AssistedManipulation(C.Structure):
    _fields_ = [("enable_joint_limit", C.c_int32), ("enable_self_collision_limit", C.c_int32),
                ("enable_workspace_limit", C.c_int32), ("enable_energy_limit", C.c_int32),
                ("enable_velocity_cost", C.c_int32), ("enable_trajectory_cost", C.c_int32),
                ("enable_manipulability_cost", C.c_int32), ("link_position_mode", C.c_int32),
                ("lower_joint_limit", Barrier * 12), ("upper_joint_limit", Barrier * 12),
                ("self_collision_limit", Barrier), ("self_collision_radii", C.c_double * 8),
                ("workspace_limit_above", Barrier), ("workspace_limit_infront", Barrier),
                ("workspace_limit_reach", Barrier), ("workspace_cost_yaw", Quadratic),
                ("energy_limit_below", Barrier), ("energy_limit_above", Barrier),
                ("velocity_cost", Quadratic * 12),
                ("trajectory_target_scale", C.c_double), ("trajectory_target_maximum", C.c_double),
                ("trajectory_position_cost", Quadratic), ("trajectory_position_threshold", C.c_double),
                ("trajectory_velocity_cost", Quadratic),
                ("trajectory_velocity_minimum", C.c_double), ("trajectory_velocity_maximum", C.c_double),
                ("trajectory_velocity_dropoff", C.c_double), ("manipulability_cost", Quadratic)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("system", C.c_int32), ("objective", C.c_int32),
                ("precision", C.c_int32), ("dynamics_mode", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("world_size", C.c_int32),
                ("state_dof", C.c_int32), ("control_dof", C.c_int32),
                ("rollouts", C.c_int64), ("keep_best_rollouts", C.c_int64),
                ("time_step", C.c_double), ("horison", C.c_double), ("gradient_step", C.c_double),
                ("cost_scale", C.c_double), ("cost_discount_factor", C.c_double),
                ("covariance", C.POINTER(C.c_double)), ("covariance_rows", C.c_int32), ("covariance_cols", C.c_int32),
                ("control_bound", C.c_int32), ("control_limits_size", C.c_int32),
                ("control_min", C.POINTER(C.c_double)), ("control_max", C.POINTER(C.c_double)),
                ("control_default", C.POINTER(C.c_double)),
                ("smoothing", C.c_int32), ("smoothing_window", C.c_uint32), ("smoothing_order", C.c_uint32),
                ("threads", C.c_int32), ("batch", C.c_int32)]


class ForecastConfig(C.Structure):
    """mppi_b200_forecast_config (reference Forecast::Configuration, forecast.hpp:388-413)."""
    _fields_ = [("type", C.c_int32), ("batch", C.c_int32), ("device", C.c_int32), ("order", C.c_uint32),
                ("time_step", C.c_double), ("horison", C.c_double), ("window", C.c_double)]


FORECAST_LOCF, FORECAST_AVERAGE, FORECAST_KALMAN = 0, 1, 2


class DynamicsForecastConfig(C.Structure):
    """mppi_b200_dynamics_forecast_config (reference DynamicsForecast::Configuration, dynamics.hpp:173-188)."""
    _fields_ = [("batch", C.c_int32), ("device", C.c_int32), ("time_step", C.c_double), ("horison", C.c_double),
                ("apply_wrench", C.c_int32), ("reserved", C.c_int32)]


DYNAMICS_FORECAST_RECORD = 112
# slices of one record (include/mppi_b200.h)
DF_JOINT_POSITION, DF_POSITION, DF_ORIENTATION = slice(0, 12), slice(12, 15), slice(15, 19)
DF_LINEAR_VELOCITY, DF_ANGULAR_VELOCITY = slice(19, 22), slice(22, 25)
DF_LINEAR_ACCELERATION, DF_ANGULAR_ACCELERATION = slice(25, 28), slice(28, 31)
DF_JOINT_POWER, DF_EXTERNAL_POWER, DF_ENERGY, DF_WRENCH, DF_JACOBIAN = 31, 32, 33, slice(34, 40), slice(40, 112)

EXPORTS = [
    "mppi_b200_create", "mppi_b200_destroy", "mppi_b200_last_error", "mppi_b200_update",
    "mppi_b200_update_begin", "mppi_b200_update_weights", "mppi_b200_update_finish",
    "mppi_b200_reduce_buffers", "mppi_b200_stream", "mppi_b200_synchronize",
    "mppi_b200_comm_unique_id", "mppi_b200_comm_init", "mppi_b200_get", "mppi_b200_read",
    "mppi_b200_query", "mppi_b200_last_update_device_seconds", "mppi_b200_default_track_point",
    "mppi_b200_default_assisted_manipulation", "mppi_b200_default_toy_objective",
    "mppi_b200_set_profiling", "mppi_b200_stage_seconds", "mppi_b200_measure_fma_peak",
    "mppi_b200_update_launch", "mppi_b200_update_wait",
    "mppi_b200_forecast_create", "mppi_b200_forecast_destroy", "mppi_b200_forecast_last_error",
    "mppi_b200_forecast_update", "mppi_b200_forecast_update_time", "mppi_b200_forecast_table",
    "mppi_b200_forecast_table_device", "mppi_b200_set_wrench_device", "mppi_b200_forecast_batch",
    "mppi_b200_dynamics_forecast_create", "mppi_b200_dynamics_forecast_destroy", "mppi_b200_dynamics_forecast_last_error",
    "mppi_b200_dynamics_forecast_steps", "mppi_b200_dynamics_forecast_run", "mppi_b200_dynamics_forecast_read",
    "mppi_b200_dynamics_forecast_device_records",
    "mppi_b200_p2p_handle", "mppi_b200_p2p_init",
]
STAGES = ("h2d", "warm_start_shift", "sample", "rollout", "weights", "weighted_sum", "finish", "d2h")

_dp = C.POINTER(C.c_double)


def library_path():
    # MPPI_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("MPPI_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libmppi_b200.so")


def load_library(path=None):
    """Load the CUDA engine's shared library. Fails loudly when it has not been built."""
    path = path or library_path()
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.mppi_b200_create.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.mppi_b200_create.restype = C.c_int
    lib.mppi_b200_destroy.argtypes = [C.c_void_p]
    lib.mppi_b200_destroy.restype = None
    lib.mppi_b200_last_error.argtypes = [C.c_void_p]
    lib.mppi_b200_last_error.restype = C.c_char_p
    upd = [C.c_void_p, _dp, C.c_double, _dp, C.c_void_p, C.c_int32, C.c_uint64]
    lib.mppi_b200_update.argtypes = upd
    lib.mppi_b200_update.restype = C.c_int
    lib.mppi_b200_update_begin.argtypes = upd
    lib.mppi_b200_update_begin.restype = C.c_int
    lib.mppi_b200_update_launch.argtypes = upd
    lib.mppi_b200_update_launch.restype = C.c_int
    lib.mppi_b200_update_wait.argtypes = [C.c_void_p]
    lib.mppi_b200_update_wait.restype = C.c_int
    for f in (lib.mppi_b200_update_weights, lib.mppi_b200_update_finish, lib.mppi_b200_synchronize):
        f.argtypes = [C.c_void_p]
        f.restype = C.c_int
    lib.mppi_b200_reduce_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.mppi_b200_reduce_buffers.restype = C.c_int
    lib.mppi_b200_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mppi_b200_stream.restype = C.c_int
    lib.mppi_b200_comm_unique_id.argtypes = [C.c_void_p]
    lib.mppi_b200_comm_unique_id.restype = C.c_int
    lib.mppi_b200_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.mppi_b200_comm_init.restype = C.c_int
    lib.mppi_b200_get.argtypes = [C.c_void_p, _dp, C.c_double]
    lib.mppi_b200_get.restype = C.c_int
    lib.mppi_b200_read.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
    lib.mppi_b200_read.restype = C.c_int
    lib.mppi_b200_query.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
    lib.mppi_b200_query.restype = C.c_int
    lib.mppi_b200_last_update_device_seconds.argtypes = [C.c_void_p, _dp]
    lib.mppi_b200_last_update_device_seconds.restype = C.c_int
    lib.mppi_b200_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    lib.mppi_b200_set_profiling.restype = C.c_int
    lib.mppi_b200_stage_seconds.argtypes = [C.c_void_p, _dp, C.c_size_t]
    lib.mppi_b200_stage_seconds.restype = C.c_int
    lib.mppi_b200_measure_fma_peak.argtypes = [C.c_int32, C.c_int32, _dp]
    lib.mppi_b200_measure_fma_peak.restype = C.c_int
    lib.mppi_b200_default_track_point.argtypes = [C.POINTER(TrackPoint)]
    lib.mppi_b200_default_assisted_manipulation.argtypes = [C.POINTER(AssistedManipulation)]
    lib.mppi_b200_default_toy_objective.argtypes = [C.POINTER(ToyObjective)]
    lib.mppi_b200_forecast_create.argtypes = [C.POINTER(ForecastConfig), _dp, C.POINTER(C.c_void_p)]
    lib.mppi_b200_forecast_create.restype = C.c_int
    lib.mppi_b200_forecast_destroy.argtypes = [C.c_void_p]
    lib.mppi_b200_forecast_destroy.restype = None
    lib.mppi_b200_forecast_last_error.argtypes = [C.c_void_p]
    lib.mppi_b200_forecast_last_error.restype = C.c_char_p
    lib.mppi_b200_forecast_update.argtypes = [C.c_void_p, _dp, C.c_double]
    lib.mppi_b200_forecast_update.restype = C.c_int
    lib.mppi_b200_forecast_update_time.argtypes = [C.c_void_p, C.c_double]
    lib.mppi_b200_forecast_update_time.restype = C.c_int
    lib.mppi_b200_forecast_table.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32, _dp]
    lib.mppi_b200_forecast_table.restype = C.c_int
    lib.mppi_b200_forecast_table_device.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_void_p)]
    lib.mppi_b200_forecast_table_device.restype = C.c_int
    lib.mppi_b200_set_wrench_device.argtypes = [C.c_void_p, C.c_void_p]
    lib.mppi_b200_set_wrench_device.restype = C.c_int
    lib.mppi_b200_forecast_batch.argtypes = [C.c_void_p]
    lib.mppi_b200_forecast_batch.restype = C.c_int
    lib.mppi_b200_dynamics_forecast_create.argtypes = [C.POINTER(DynamicsForecastConfig), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mppi_b200_dynamics_forecast_create.restype = C.c_int
    lib.mppi_b200_dynamics_forecast_destroy.argtypes = [C.c_void_p]
    lib.mppi_b200_dynamics_forecast_destroy.restype = None
    lib.mppi_b200_dynamics_forecast_last_error.argtypes = [C.c_void_p]
    lib.mppi_b200_dynamics_forecast_last_error.restype = C.c_char_p
    lib.mppi_b200_dynamics_forecast_steps.argtypes = [C.c_void_p]
    lib.mppi_b200_dynamics_forecast_steps.restype = C.c_int
    lib.mppi_b200_dynamics_forecast_run.argtypes = [C.c_void_p, _dp, C.c_double]
    lib.mppi_b200_dynamics_forecast_run.restype = C.c_int
    lib.mppi_b200_dynamics_forecast_read.argtypes = [C.c_void_p, _dp, C.c_size_t]
    lib.mppi_b200_dynamics_forecast_read.restype = C.c_int
    lib.mppi_b200_dynamics_forecast_device_records.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mppi_b200_dynamics_forecast_device_records.restype = C.c_int
    lib.mppi_b200_p2p_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.mppi_b200_p2p_handle.restype = C.c_int
    lib.mppi_b200_p2p_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.mppi_b200_p2p_init.restype = C.c_int
    return lib


# ---- reference defaults (host data only) -----------------------------------------------------

def connect_ranks(lib, engine_handle, dist, torch, exchange="p2p"):
    """Attach the in-library exchange of a sharded engine (one process per GPU). `p2p`: every rank's mailbox handle is
    all-gathered through torch.distributed and opened by its peers; `nccl`: rank 0's NCCL id is broadcast."""
    rank = dist.get_rank()
    if exchange == "p2p":
        buf = (C.c_ubyte * 64)()
        rc = lib.mppi_b200_p2p_handle(engine_handle, buf)
        mine = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")   # every rank takes part in the gather, also one that failed
        every = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(dist.get_world_size())]
        dist.all_gather(every, mine)
        if rc != 0:
            return rc
        raw = bytes(torch.cat(every).cpu().tolist())
        return lib.mppi_b200_p2p_init(engine_handle, C.c_char_p(raw))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        rc = lib.mppi_b200_comm_unique_id(buf)
        if rc != 0:
            return rc
        uid.copy_(torch.tensor(list(buf), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    return lib.mppi_b200_comm_init(engine_handle, C.c_char_p(bytes(uid.cpu().tolist())))


def huddled_state(energy=100.0):
    """make_state(Preset::HUDDLED), reference src/frankaridgeback/state.cpp:15-19."""
    x = np.zeros(31)
    x[:12] = [0.2, 0.2, math.pi / 4, 0.0, math.pi / 5, 0.0, -math.pi / 2, 0.0, 2, math.pi / 4, 0.025, 0.025]
    x[30] = energy
    return x


FRANKA_COVARIANCE_DIAG = np.array([0.1, 0.1, 0.2] + [7.5] * 7 + [0.0, 0.0])  # base.hpp:79-83
FRANKA_CONTROL_MIN = np.array([-0.5, -0.5, -1.0] + [-100.0] * 7 + [-0.05, -0.05])  # base.hpp:85-89
FRANKA_CONTROL_MAX = -FRANKA_CONTROL_MIN  # base.hpp:90-94


class ConfigHolder:
    """Owns the numpy arrays a Config points into."""

    def __init__(self, cfg, keep):
        self.cfg, self._keep = cfg, keep


def make_config(system, objective, rollouts, horison, *, precision=FP64, dynamics_mode=DYNAMICS_FAITHFUL,
                keep_best=0, time_step=0.01, gradient_step=2.0, cost_scale=10.0, discount=1.0,
                covariance=None, control_min=None, control_max=None, control_bound=True,
                control_default=None, smoothing=(10, 1), threads=1, device=0, rank=0, world_size=1, batch=1):
    """mppi::Configuration with the defaults of src/test/case/base.hpp:68-100."""
    if system == SYSTEM_TOY:
        nx, nu = 4, 2
        covariance = np.eye(2) if covariance is None else covariance
        control_min = np.full(2, -5.0) if control_min is None else control_min
        control_max = np.full(2, 5.0) if control_max is None else control_max
    else:
        nx, nu = 31, 12
        covariance = np.diag(FRANKA_COVARIANCE_DIAG) if covariance is None else covariance
        control_min = FRANKA_CONTROL_MIN if control_min is None else control_min
        control_max = FRANKA_CONTROL_MAX if control_max is None else control_max
    cov = np.asfortranarray(np.array(covariance, dtype=np.float64))
    cmin = np.ascontiguousarray(control_min, dtype=np.float64)
    cmax = np.ascontiguousarray(control_max, dtype=np.float64)
    cdef = None if control_default is None else np.ascontiguousarray(control_default, dtype=np.float64)
    c = Config()
    c.abi_version = ABI_VERSION
    c.system, c.objective, c.precision, c.dynamics_mode = system, objective, precision, dynamics_mode
    c.device, c.rank, c.world_size = device, rank, world_size
    c.state_dof, c.control_dof = nx, nu
    c.rollouts, c.keep_best_rollouts = rollouts, keep_best
    c.time_step, c.horison, c.gradient_step = time_step, horison, gradient_step
    c.cost_scale, c.cost_discount_factor = cost_scale, discount
    c.covariance = cov.ctypes.data_as(_dp)
    c.covariance_rows, c.covariance_cols = cov.shape
    c.control_bound = int(control_bound)
    c.control_limits_size = len(cmin)
    c.control_min = cmin.ctypes.data_as(_dp)
    c.control_max = cmax.ctypes.data_as(_dp)
    c.control_default = cdef.ctypes.data_as(_dp) if cdef is not None else None
    c.smoothing = int(smoothing is not None)
    c.smoothing_window, c.smoothing_order = smoothing if smoothing is not None else (0, 0)
    c.threads = threads
    c.batch = batch
    return ConfigHolder(c, (cov, cmin, cmax, cdef))


def default_toy_objective():
    """BASELINE.json config 1: 100|p-(1,1)|^2 + |v|^2 + 0.01|u|^2."""
    o = ToyObjective()
    o.target[0], o.target[1] = 1.0, 1.0
    o.position_cost, o.velocity_cost, o.control_cost = 100.0, 1.0, 0.01
    return o


_ARM_LOWER = [-2.0, -2.0, -6.28, -2.8, -1.745, -2.8, -3.0718, -2.7925, 0.349, -2.967, 0.0, 0.0]
_ARM_UPPER = [2.0, 2.0, 6.28, 2.8, 1.745, 2.8, 0.0, 2.7925, 4.53785, 2.967, 0.5, 0.5]


def default_track_point():
    """TrackPoint::DEFAULT_CONFIGURATION, objective/track_point.hpp:77-114."""
    o = TrackPoint()
    o.point[:] = [1.0, 1.0, 1.0]
    o.enable_joint_limits, o.enable_self_collision_avoidance, o.enable_power_limit, o.enable_reach_limits = 1, 0, 0, 0
    lo_s = [1.0, 0.0, 0.0, 10.0, 50.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0]
    hi_s = [0.0, 0.0, 0.0, 10.0, 50.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 10.0]
    for i in range(12):
        o.lower_joint_limit[i] = Barrier(_ARM_LOWER[i], lo_s[i], 1e10)
        o.upper_joint_limit[i] = Barrier(_ARM_UPPER[i], hi_s[i], 1e10)
    o.self_collision_limit = Barrier(0.0, 1.0, 1e10)
    o.self_collision_radii[:] = [0.75] + [0.1] * 7
    o.maximum_reach_limit = Barrier(0.8, 1.0, 1e10)
    o.link_position_mode = LINKS_ZERO
    return o


def default_assisted_manipulation():
    """AssistedManipulation::DEFAULT_CONFIGURATION, objective/assisted_manipulation.hpp:133-206."""
    o = AssistedManipulation()
    (o.enable_joint_limit, o.enable_self_collision_limit, o.enable_workspace_limit, o.enable_energy_limit,
     o.enable_velocity_cost, o.enable_trajectory_cost, o.enable_manipulability_cost) = 1, 1, 1, 0, 1, 1, 1
    o.link_position_mode = LINKS_ZERO
    s = [0.0, 0.0, 0.0] + [10.0] * 7 + [0.0, 0.0]
    for i in range(12):
        o.lower_joint_limit[i] = Barrier(_ARM_LOWER[i], s[i], 1e10)
        o.upper_joint_limit[i] = Barrier(_ARM_UPPER[i], s[i], 1e10)
    o.self_collision_limit = Barrier(0.0, 1.0, 1e10)
    o.self_collision_radii[:] = [0.75] + [0.1] * 7
    o.workspace_limit_above = Barrier(0.0, 1.0, 1e10)
    o.workspace_limit_infront = Barrier(0.0, 1.0, 1e10)
    o.workspace_limit_reach = Barrier(1.0, 1.0, 1e10)
    o.workspace_cost_yaw = Quadratic(0.0, 0.0, 400.0)
    o.energy_limit_below = Barrier(0.0, 10.0, 1e10)
    o.energy_limit_above = Barrier(20.0, 10.0, 1e10)
    for i, w in enumerate([1000.0, 1000.0, 100.0, 0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.0, 0.0]):
        o.velocity_cost[i] = Quadratic(0.0, 0.0, w)
    o.trajectory_target_scale, o.trajectory_target_maximum = 1e-2, 1.0
    o.trajectory_position_cost = Quadratic(100.0, 0.0, 500.0)
    o.trajectory_position_threshold = 0.0
    o.trajectory_velocity_cost = Quadratic(0.0, 0.0, 500.0)
    o.trajectory_velocity_minimum, o.trajectory_velocity_maximum, o.trajectory_velocity_dropoff = 0.1, 5.0, 2.0
    o.manipulability_cost = Quadratic(0.0, 0.0, 10.0)
    return o


def philox_like_noise(seed, shape):
    """Host noise for injected-noise parity runs: counter-based, reproducible (numpy Philox)."""
    return np.random.Generator(np.random.Philox(key=seed)).standard_normal(shape)
