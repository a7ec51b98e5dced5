"""SURVEY §8f-4, CPU: the on-disk formats (assistedmanipulation_b200/formats.py) against files written by the
reference's own logger over its own Trajectory (tests/golden/ref_logs/, tools/gen_log_golden.py), and the JSON
layout of the reference's configuration structs."""
import json
import os

import numpy as np
import pytest

import cases
import oracle_lib as ol
from assistedmanipulation_b200 import abi, formats

GOLDEN = os.path.join(ol.ROOT, "tests", "golden", "ref_logs")


@pytest.mark.parametrize("name", sorted(cases.LOG_CASES))
def test_csv_logs_are_byte_identical_with_the_reference_logger(oracle, tmp_path, name):
    """The oracle (bit exact with the reference build, own mt19937 sampling) feeds formats.MPPILog; every file
    must equal what logger::MPPI wrote for the reference's Trajectory, byte for byte (update.csv: all but the
    measured duration)."""
    case = cases.LOG_CASES[name]
    holder = cases.config_for(case)
    o = ol.Oracle(oracle, holder, case["params"]())
    log = formats.MPPILog(str(tmp_path), holder.cfg.control_dof, case["K"] + 2)
    for u in range(case["updates"]):
        t = u * case["cadence"]
        assert o.update(case["x0"], t, None, None) == 0
        log.log(o, t, holder.cfg.time_step, update_duration=0.0)
        log.log(o, t, holder.cfg.time_step)   # same update time: nothing written (mppi.cpp:86-88)
    log.close()
    o.close()
    for f in formats.MPPILog.FILES:
        got = open(os.path.join(str(tmp_path), f + ".csv")).read()
        want = open(os.path.join(GOLDEN, name, f + ".csv")).read()
        if f == "update":
            strip = lambda text: [line.rsplit(", ", 1)[0] for line in text.splitlines()[1:]]
            assert got.splitlines()[0] == want.splitlines()[0] and strip(got) == strip(want)
        else:
            assert got == want, f


def test_number_formatting_is_ostream_default():
    # operator<<(double): %g with 6 significant digits
    for x, text in [(0.0, "0"), (1.0, "1"), (0.02, "0.02"), (1120.0812, "1120.08"), (1e-5, "1e-05"), (123456789.0, "1.23457e+08"),
                    (2e11 + 4.2125, "2e+11"), (-0.5, "-0.5"), (float("inf"), "inf"), (float("nan"), "nan")]:
        assert formats.format_double(x) == text
    assert formats.format_value(np.int64(3)) == "3" and formats.format_value([1.5, 2]) == "1.5, 2"


def test_csv_round_trip(tmp_path):
    path = str(tmp_path / "deep" / "folder" / "a.csv")   # parent folders are created (file.hpp:27-40)
    c = formats.CSV(path, formats.CSV.make_header("update", "time", ["x1", "x2"]))
    c.write(1, 0.5, np.array([1.25, -2.0]))
    c.close()
    assert open(path).read() == "update, time, x1, x2\n1, 0.5, 1.25, -2\n"
    header, rows = formats.read_csv(path)
    assert header == ["update", "time", "x1", "x2"] and np.array_equal(rows, [[1, 0.5, 1.25, -2.0]])


def test_merge_patch_rfc7386():
    # the examples of RFC 7386 section 3 / appendix A that nlohmann::json::merge_patch implements
    t = {"title": "Goodbye!", "author": {"givenName": "John", "familyName": "Doe"}, "tags": ["example", "sample"], "content": "x"}
    p = {"title": "Hello!", "phoneNumber": "+01", "author": {"familyName": None}, "tags": ["example"]}
    assert formats.merge_patch(t, p) == {"title": "Hello!", "author": {"givenName": "John"}, "tags": ["example"], "content": "x", "phoneNumber": "+01"}
    assert formats.merge_patch({"a": "b"}, {"a": None}) == {}
    assert formats.merge_patch({"a": [{"b": "c"}]}, {"a": [1]}) == {"a": [1]}
    assert formats.merge_patch({"e": None}, {"a": 1}) == {"e": None, "a": 1}
    assert formats.merge_patch([1, 2], {"a": "b", "c": None}) == {"a": "b"}


def test_mppi_configuration_json_round_trip():
    holder = abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, 4096, 0.64, keep_best=20, smoothing=(10, 1), threads=12,
                             control_default=np.zeros(12))
    j = formats.mppi_configuration_to_json(holder)
    assert list(j) == ["initial_state", "rollouts", "keep_best_rollouts", "time_step", "horison", "gradient_step", "cost_scale",
                       "cost_discount_factor", "covariance", "control_bound", "control_min", "control_max", "control_default", "smoothing", "threads"]  # mppi.hpp:243-248
    assert j["control_min"][3] == [-100.0] and len(j["covariance"]) == 12 and len(j["covariance"][0]) == 12     # json.hpp:49-63: rows of a column vector
    assert j["smoothing"] == {"window": 10, "order": 1} and len(j["initial_state"]) == 31
    text = json.dumps(j)
    # the CLI's merge patch (base.cpp:12-24): change the rollout count, drop smoothing and the default control
    patched = formats.merge_patch(json.loads(text), {"rollouts": 512, "smoothing": {"window": None, "order": None}, "control_default": {}})
    back = formats.mppi_configuration_from_json(patched, abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, precision=abi.FP32, batch=4)
    c = back.cfg
    assert (c.rollouts, c.keep_best_rollouts, c.smoothing, c.threads, c.precision, c.batch) == (512, 20, 0, 12, abi.FP32, 4)
    assert c.horison == 0.64 and not c.control_default                                                             # std::optional <-> {} (json.hpp:15-34)
    again = formats.mppi_configuration_to_json(back)
    assert again["smoothing"] == {} and again["control_default"] == {} and again["covariance"] == j["covariance"]


def test_assisted_manipulation_json_round_trip():
    p = abi.default_assisted_manipulation()
    j = formats.assisted_manipulation_to_json(p)
    assert list(j)[:9] == ["enable_joint_limit", "enable_self_collision_limit", "enable_workspace_limit", "enable_energy_limit", "enable_velocity_cost",
                           "enable_trajectory_cost", "enable_manipulability_cost", "lower_joint_limit", "upper_joint_limit"]   # assisted_manipulation.hpp:95-125
    assert set(j["lower_joint_limit"][0]) == {"lower_bound", "scale", "maximum_cost"} and set(j["upper_joint_limit"][0]) == {"upper_bound", "scale", "maximum_cost"}
    assert set(j["velocity_cost"][0]) == {"linear_cost", "constant_cost", "quadratic_cost"} and "self_collision_radii" not in j
    patched = formats.merge_patch(json.loads(json.dumps(j)), {"enable_energy_limit": False, "energy_limit_above": {"upper_bound": 25.0},
                                                              "trajectory_velocity_dropoff": 3.0})
    q = formats.assisted_manipulation_from_json(patched)
    assert q.enable_energy_limit == 0 and q.energy_limit_above.bound == 25.0 and q.energy_limit_above.scale == p.energy_limit_above.scale
    assert q.trajectory_velocity_dropoff == 3.0
    assert formats.assisted_manipulation_to_json(formats.assisted_manipulation_from_json(j)) == j


def test_track_point_json_round_trip():
    p = abi.default_track_point()
    j = formats.track_point_to_json(p)
    assert list(j) == ["point", "enable_joint_limits", "enable_self_collision_avoidance", "enable_power_limit", "enable_reach_limits",
                       "lower_joint_limit", "upper_joint_limit", "self_collision_limit", "self_collision_radii", "maximum_reach_limit"]   # track_point.hpp:50-56
    assert j["point"] == [[1.0], [1.0], [1.0]] and len(j["self_collision_radii"]) == 8
    q = formats.track_point_from_json(formats.merge_patch(json.loads(json.dumps(j)), {"point": [[0.5], [0.0], [1.2]], "enable_reach_limits": True}))
    assert list(q.point) == [0.5, 0.0, 1.2] and q.enable_reach_limits == 1 and q.enable_joint_limits == p.enable_joint_limits
    assert formats.track_point_to_json(formats.track_point_from_json(j)) == j


def test_forecast_configuration_json():
    cfg = abi.ForecastConfig(type=abi.FORECAST_KALMAN, batch=1, device=0, order=1, time_step=0.01, horison=1.0, window=0.0)
    j = formats.forecast_configuration_to_json(cfg)
    assert list(j) == ["type", "locf", "average", "kalman"] and j["type"] == 2 and j["locf"] == {} and j["average"] == {}   # forecast.hpp:413-416
    assert set(j["kalman"]) == {"observed_states", "time_step", "horison", "order", "variance", "initial_state"}
    back, initial = formats.forecast_configuration_from_json(json.loads(json.dumps(j)), batch=4, device=1)
    assert (back.type, back.order, back.time_step, back.horison, back.batch, back.device) == (2, 1, 0.01, 1.0, 4, 1) and np.array_equal(initial, np.zeros(6))
    locf = formats.forecast_configuration_to_json(abi.ForecastConfig(type=abi.FORECAST_LOCF, batch=1, device=0, order=0, time_step=0, horison=0.4, window=0), observation=[1, 2, 3, 0, 0, 0])
    back, initial = formats.forecast_configuration_from_json(locf)
    assert back.horison == 0.4 and list(initial) == [1, 2, 3, 0, 0, 0]
    with pytest.raises(AssertionError, match="no configuration provided"):
        formats.forecast_configuration_from_json({"type": 1, "locf": {}, "average": {}, "kalman": {}})
