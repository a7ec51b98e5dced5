"""On-disk formats of the reference either side of the MPPI path, host side only:

* the CSV files `logger::MPPI` writes (reference src/logging/mppi.cpp:9-136 over logging/csv.hpp:64-170,
  logging/file.hpp): `costs.csv`, `weights.csv`, `gradient.csv`, `optimal_rollout.csv`, `optimal_cost.csv`,
  `update.csv` — header `update, time, …`, separator `", "`, numbers as a C++ `ostream << double` prints them
  (`%g`, 6 significant digits) so archived reference runs and new runs read with the same scripts;
* the JSON layout nlohmann produces for `mppi::Configuration` (controller/mppi.hpp:243-248) and the objective
  configurations (objective/assisted_manipulation.hpp:95-125, controller/cost.hpp:33-167), matrices as arrays of
  rows (controller/json.hpp:49-81; a VectorXd is a column: `[[a], [b]]`), `std::optional` as `{}` when empty
  (json.hpp:15-34), and RFC 7386 merge patches as the test CLI applies them (test/case/base.cpp:12-24).
"""
import copy
import math
import os

import numpy as np

from . import abi

SEPARATOR = ", "


# ---- numbers as `std::ostream << x` prints them ------------------------------------------------------------

def format_double(x):
    """operator<<(double) with the default precision 6 and floatfield (== printf("%g"))."""
    x = float(x)
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    if math.isinf(x):
        return "-inf" if x < 0 else "inf"
    return "%g" % x


def format_value(v):
    if isinstance(v, (bool, np.bool_)):
        return "1" if v else "0"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return format_double(v)
    return SEPARATOR.join(format_value(e) for e in np.asarray(v).ravel())   # iterable: every element (csv.hpp:150-163)


class CSV:
    """logger::CSV (logging/csv.hpp): header on creation, one `write` per row, parent folders created."""

    def __init__(self, path, header):
        parent = os.path.dirname(path)
        if parent:
            os.makedirs(parent, exist_ok=True)
        self.file = open(path, "w", newline="")
        if header:
            self.file.write(SEPARATOR.join(header) + "\n")

    @staticmethod
    def make_header(*args):
        header = []
        for a in args:
            header.extend([a] if isinstance(a, str) else list(a))
        return header

    def write(self, *args):
        self.file.write(SEPARATOR.join(format_value(a) for a in args) + "\n")

    def flush(self):
        self.file.flush()

    def close(self):
        self.file.close()


def read_csv(path):
    """-> (header list, float array). Reads what CSV / logger::CSV write."""
    with open(path) as f:
        header = f.readline().rstrip("\n").split(SEPARATOR)
        rows = [[float(c) for c in line.rstrip("\n").split(SEPARATOR)] for line in f if line.strip()]
    return header, np.array(rows)


class MPPILog:
    """logger::MPPI (logging/mppi.cpp). `source` is anything with the engine's call shapes
    (engine.Engine, oracle_lib.Oracle): read(what, count), query(what)."""

    FILES = ("costs", "weights", "gradient", "optimal_rollout", "optimal_cost", "update")

    def __init__(self, folder, control_dof, rollouts, log_costs=True, log_weights=True, log_gradient=True, log_optimal_rollout=True,
                 log_optimal_cost=True, log_update=True):
        control = ["control%d" % i for i in range(1, control_dof + 1)]
        rollout = ["rollout%d" % i for i in range(1, rollouts + 1)]
        mk = lambda name, *cols: CSV(os.path.join(folder, name + ".csv"), CSV.make_header("update", "time", *cols))
        self.costs = mk("costs", rollout) if log_costs else None
        self.weights = mk("weights", rollout) if log_weights else None
        self.gradient = mk("gradient", control) if log_gradient else None
        self.optimal_rollout = mk("optimal_rollout", control) if log_optimal_rollout else None
        self.optimal_cost = mk("optimal_cost", "cost") if log_optimal_cost else None
        self.update = mk("update", "update_duration") if log_update else None
        self.last_update = None

    def log(self, source, time, time_step, update_duration=0.0):
        """One row set per controller update (mppi.cpp:84-136); `time` = Trajectory::get_update_last()."""
        if time == self.last_update:
            return
        steps, rollouts = source.query(abi.QUERY_STEP_COUNT), source.query(abi.QUERY_ROLLOUT_COUNT)
        nu = source.query(abi.QUERY_CONTROL_DOF)
        iteration = source.query(abi.QUERY_UPDATE_COUNT)
        if self.update:
            self.update.write(iteration, float(time), float(update_duration))
        times = [time + i * time_step for i in range(steps)]
        if self.costs:
            self.costs.write(iteration, float(time), source.read(abi.READ_COSTS, rollouts))
        if self.weights:
            self.weights.write(iteration, float(time), source.read(abi.READ_WEIGHTS, rollouts))
        if self.gradient:
            g = source.read(abi.READ_GRADIENT, nu * steps).reshape(steps, nu)   # device layout [T][nu] = the columns of the nu x T matrix
            for i in range(steps):
                self.gradient.write(iteration, float(times[i]), g[i])
        if self.optimal_rollout:
            u = source.read(abi.READ_OPTIMAL, nu * steps).reshape(steps, nu)
            for i in range(steps):
                self.optimal_rollout.write(iteration, float(times[i]), u[i])
        if self.optimal_cost:
            self.optimal_cost.write(iteration, float(time), float(source.read(abi.READ_OPTIMAL_COST, 1)[0]))
        self.last_update = time

    def close(self):
        for name in self.FILES:
            f = getattr(self, name)
            if f:
                f.close()


# ---- JSON (nlohmann layout) ----------------------------------------------------------------------------------

def merge_patch(target, patch):
    """RFC 7386, as nlohmann::json::merge_patch (test/case/base.cpp:15-19)."""
    if not isinstance(patch, dict):
        return copy.deepcopy(patch)
    if not isinstance(target, dict):
        target = {}
    out = dict(target)
    for k, v in patch.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = merge_patch(out.get(k), v)
    return out


def matrix_to_json(m):
    m = np.asarray(m, dtype=np.float64)
    if m.ndim == 1:
        m = m.reshape(-1, 1)   # a VectorXd is an n x 1 matrix
    return [[float(x) for x in row] for row in m]


def matrix_from_json(j):
    if not j:
        return np.zeros((0, 0))
    return np.array(j, dtype=np.float64).reshape(len(j), len(j[0]))


def vector_from_json(j):
    return matrix_from_json(j).reshape(-1)


def mppi_configuration_to_json(holder, initial_state=None):
    """mppi::Configuration -> the dict nlohmann would dump (field order of controller/mppi.hpp:243-248)."""
    c = holder.cfg
    nu = c.control_dof
    cov = np.ctypeslib.as_array(c.covariance, shape=(c.covariance_rows * c.covariance_cols,)).reshape(c.covariance_cols, c.covariance_rows).T
    lim = c.control_limits_size
    return {
        "initial_state": matrix_to_json(np.zeros(c.state_dof) if initial_state is None else initial_state),
        "rollouts": int(c.rollouts), "keep_best_rollouts": int(c.keep_best_rollouts),
        "time_step": c.time_step, "horison": c.horison, "gradient_step": c.gradient_step,
        "cost_scale": c.cost_scale, "cost_discount_factor": c.cost_discount_factor,
        "covariance": matrix_to_json(cov),
        "control_bound": bool(c.control_bound),
        "control_min": matrix_to_json(np.ctypeslib.as_array(c.control_min, shape=(lim,))) if lim else [],
        "control_max": matrix_to_json(np.ctypeslib.as_array(c.control_max, shape=(lim,))) if lim else [],
        "control_default": matrix_to_json(np.ctypeslib.as_array(c.control_default, shape=(nu,))) if c.control_default else {},
        "smoothing": {"window": int(c.smoothing_window), "order": int(c.smoothing_order)} if c.smoothing else {},
        "threads": int(c.threads),
    }


def mppi_configuration_from_json(j, system, objective, **engine_options):
    """The reference's JSON (possibly merge-patched) -> abi.ConfigHolder. engine_options: precision,
    dynamics_mode, device, rank, world_size, batch (what the reference has no field for)."""
    cov = matrix_from_json(j["covariance"])
    smoothing = j.get("smoothing") or None
    default = j.get("control_default")
    kw = dict(keep_best=int(j["keep_best_rollouts"]), time_step=float(j["time_step"]), gradient_step=float(j["gradient_step"]),
              cost_scale=float(j["cost_scale"]), discount=float(j["cost_discount_factor"]), covariance=cov,
              control_bound=bool(j["control_bound"]), control_min=vector_from_json(j["control_min"]), control_max=vector_from_json(j["control_max"]),
              control_default=vector_from_json(default) if default else None,
              smoothing=(int(smoothing["window"]), int(smoothing["order"])) if smoothing else None, threads=int(j.get("threads", 1)))
    kw.update(engine_options)
    return abi.make_config(system, objective, int(j["rollouts"]), float(j["horison"]), **kw)


def _barrier_to_json(b, lower):
    return {("lower_bound" if lower else "upper_bound"): b.bound, "scale": b.scale, "maximum_cost": b.maximum_cost}


def _barrier_from_json(j, out):
    out.bound = float(j["lower_bound"] if "lower_bound" in j else j["upper_bound"])
    out.scale, out.maximum_cost = float(j["scale"]), float(j["maximum_cost"])


def _quadratic_to_json(q):
    return {"linear_cost": q.linear_cost, "constant_cost": q.constant_cost, "quadratic_cost": q.quadratic_cost}


def _quadratic_from_json(j, out):
    out.linear_cost, out.constant_cost, out.quadratic_cost = float(j["linear_cost"]), float(j["constant_cost"]), float(j["quadratic_cost"])


_AM_FLAGS = ("enable_joint_limit", "enable_self_collision_limit", "enable_workspace_limit", "enable_energy_limit", "enable_velocity_cost",
             "enable_trajectory_cost", "enable_manipulability_cost")
_AM_LOWER = ("self_collision_limit", "workspace_limit_above", "workspace_limit_infront", "energy_limit_below")
_AM_UPPER = ("workspace_limit_reach", "energy_limit_above")
_AM_QUADRATIC = ("workspace_cost_yaw", "trajectory_position_cost", "trajectory_velocity_cost", "manipulability_cost")
_AM_SCALARS = ("trajectory_target_scale", "trajectory_target_maximum", "trajectory_position_threshold", "trajectory_velocity_minimum",
               "trajectory_velocity_maximum", "trajectory_velocity_dropoff")


def assisted_manipulation_to_json(p):
    """AssistedManipulation::Configuration in the field order of assisted_manipulation.hpp:95-125
    (self_collision_radii and the engine-only link_position_mode are not part of the reference's JSON)."""
    j = {k: bool(getattr(p, k)) for k in _AM_FLAGS}
    j["lower_joint_limit"] = [_barrier_to_json(b, True) for b in p.lower_joint_limit]
    j["upper_joint_limit"] = [_barrier_to_json(b, False) for b in p.upper_joint_limit]
    for k in ("self_collision_limit", "workspace_limit_above", "workspace_limit_infront"):
        j[k] = _barrier_to_json(getattr(p, k), True)
    j["workspace_limit_reach"] = _barrier_to_json(p.workspace_limit_reach, False)
    j["workspace_cost_yaw"] = _quadratic_to_json(p.workspace_cost_yaw)
    j["energy_limit_below"] = _barrier_to_json(p.energy_limit_below, True)
    j["energy_limit_above"] = _barrier_to_json(p.energy_limit_above, False)
    j["velocity_cost"] = [_quadratic_to_json(q) for q in p.velocity_cost]
    j["trajectory_target_scale"], j["trajectory_target_maximum"] = p.trajectory_target_scale, p.trajectory_target_maximum
    j["trajectory_position_cost"] = _quadratic_to_json(p.trajectory_position_cost)
    j["trajectory_position_threshold"] = p.trajectory_position_threshold
    j["trajectory_velocity_cost"] = _quadratic_to_json(p.trajectory_velocity_cost)
    for k in ("trajectory_velocity_minimum", "trajectory_velocity_maximum", "trajectory_velocity_dropoff"):
        j[k] = getattr(p, k)
    j["manipulability_cost"] = _quadratic_to_json(p.manipulability_cost)
    return j


def assisted_manipulation_from_json(j, base=None):
    """-> abi.AssistedManipulation; fields missing from `j` keep the reference defaults (or `base`)."""
    p = base if base is not None else abi.default_assisted_manipulation()
    for k in _AM_FLAGS:
        if k in j:
            setattr(p, k, int(bool(j[k])))
    for k, lower in (("lower_joint_limit", True), ("upper_joint_limit", False)):
        if k in j:
            assert len(j[k]) == 12, k
            for i, b in enumerate(j[k]):
                _barrier_from_json(b, getattr(p, k)[i])
    for k in _AM_LOWER + _AM_UPPER:
        if k in j:
            _barrier_from_json(j[k], getattr(p, k))
    for k in _AM_QUADRATIC:
        if k in j:
            _quadratic_from_json(j[k], getattr(p, k))
    if "velocity_cost" in j:
        assert len(j["velocity_cost"]) == 12
        for i, q in enumerate(j["velocity_cost"]):
            _quadratic_from_json(q, p.velocity_cost[i])
    for k in _AM_SCALARS:
        if k in j:
            setattr(p, k, float(j[k]))
    return p


_TP_FLAGS = ("enable_joint_limits", "enable_self_collision_avoidance", "enable_power_limit", "enable_reach_limits")


def track_point_to_json(p):
    """TrackPoint::Configuration in the field order of track_point.hpp:50-56."""
    j = {"point": matrix_to_json(list(p.point))}
    for k in _TP_FLAGS:
        j[k] = bool(getattr(p, k))
    j["lower_joint_limit"] = [_barrier_to_json(b, True) for b in p.lower_joint_limit]
    j["upper_joint_limit"] = [_barrier_to_json(b, False) for b in p.upper_joint_limit]
    j["self_collision_limit"] = _barrier_to_json(p.self_collision_limit, True)
    j["self_collision_radii"] = [float(r) for r in p.self_collision_radii]
    j["maximum_reach_limit"] = _barrier_to_json(p.maximum_reach_limit, False)
    return j


def track_point_from_json(j, base=None):
    p = base if base is not None else abi.default_track_point()
    if "point" in j:
        for i, v in enumerate(vector_from_json(j["point"])):
            p.point[i] = float(v)
    for k in _TP_FLAGS:
        if k in j:
            setattr(p, k, int(bool(j[k])))
    for k in ("lower_joint_limit", "upper_joint_limit"):
        if k in j:
            assert len(j[k]) == 12, k
            for i, b in enumerate(j[k]):
                _barrier_from_json(b, getattr(p, k)[i])
    for k in ("self_collision_limit", "maximum_reach_limit"):
        if k in j:
            _barrier_from_json(j[k], getattr(p, k))
    if "self_collision_radii" in j:
        assert len(j["self_collision_radii"]) == 8
        for i, r in enumerate(j["self_collision_radii"]):
            p.self_collision_radii[i] = float(r)
    return p


def forecast_configuration_to_json(cfg, initial=None, observation=None):
    """Forecast::Configuration (forecast.hpp:391-416): `type` is a plain enum (an integer in the JSON), the three
    optional sub-configurations are `{}` when absent (controller/json.hpp:15-34)."""
    j = {"type": int(cfg.type), "locf": {}, "average": {}, "kalman": {}}
    if cfg.type == abi.FORECAST_LOCF:
        j["locf"] = {"observation": matrix_to_json(np.zeros(6) if observation is None else observation), "horison": cfg.horison}
    elif cfg.type == abi.FORECAST_AVERAGE:
        j["average"] = {"states": 6, "window": cfg.window}
    else:
        j["kalman"] = {"observed_states": 6, "time_step": cfg.time_step, "horison": cfg.horison, "order": int(cfg.order),
                       "variance": matrix_to_json(np.zeros(6)), "initial_state": matrix_to_json(np.zeros(6) if initial is None else initial)}
    return j


def forecast_configuration_from_json(j, batch=1, device=0):
    """-> (abi.ForecastConfig, initial 6-vector or None) for mppi_b200_forecast_create."""
    typ = int(j["type"])
    cfg = abi.ForecastConfig(type=typ, batch=batch, device=device, order=0, time_step=0.0, horison=0.0, window=0.0)
    initial = None
    if typ == abi.FORECAST_LOCF:
        sub = j.get("locf") or None
        assert sub, "locf forecast selected with no configuration provided"        # forecast.cpp:10-13
        cfg.horison = float(sub["horison"])
        initial = vector_from_json(sub["observation"])
    elif typ == abi.FORECAST_AVERAGE:
        sub = j.get("average") or None
        assert sub, "average forecast selected with no configuration provided"
        assert int(sub["states"]) == 6, "the device producer forecasts the 6-component wrench"
        cfg.window = float(sub["window"])
    elif typ == abi.FORECAST_KALMAN:
        sub = j.get("kalman") or None
        assert sub, "kalman forecast selected with no configuration provided"
        assert int(sub["observed_states"]) == 6, "the device producer forecasts the 6-component wrench"
        cfg.time_step, cfg.horison, cfg.order = float(sub["time_step"]), float(sub["horison"]), int(sub["order"])
        init = vector_from_json(sub.get("initial_state") or [])
        initial = init if init.size else None
    else:
        raise ValueError("unknown forecast type %d" % typ)
    return cfg, initial
