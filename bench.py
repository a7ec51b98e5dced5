#!/usr/bin/env python3
"""bench.py — the MPPI rollout path on N B200s (driver contract, tier ④).

A "step" is one mppi::Trajectory::update() (reference src/controller/mppi.cpp:154-187): sample ->
K+2 rollouts of T steps -> exp-weighting -> weighted-sum update -> smoothing/clamp. Headline workload at N=1 is
BASELINE.json configs[1]: Franka Research 3 + Ridgeback, TrackPoint objective, K=4096 x T=64, FP64.
For N>1 the rollout set grows with N (K = 4096*N, "weak") and is sharded over the ranks; the two
exchanges of the path (min/max of the costs, weighted sums) run in the library's own kernels over NVLink peer memory.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg4_f32|cfg4_f64|cfg5|toy]
                  [--no-secondary]

The same JSON line carries a `secondary` block (NOT part of `value`): a few updates of the other BASELINE.json
configurations (3, 4 in both precisions, 5) with stage times and rooflines, and for N>1 the strong-scaling config 4,
the batched config 5 and a `sharded_parity` record (the sharded engine against the same rollout set on one GPU).

`--impl reference` times the reference's CPU algorithm (the oracle port with its thread pool on all
host cores; the reference itself cannot run K > 253 nor be built whole offline) on the same
workload, one bounded sample per step.
"""
import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from assistedmanipulation_b200 import abi  # noqa: E402

# Algorithmic work per rollout-step: Featherstone operation counts at n = 12 one-dof
# joints — ABA 4641 + RNEA 1880 + second-order FK / frames / WORLD jacobian ~2300 + cost + Euler + tank.
FLOPS_PER_STEP = {"cfg2": 9000.0, "cfg3": 9500.0, "toy": 25.0, "cfg4_f32": 9000.0, "cfg4_f64": 9000.0, "cfg5": 9500.0}
COUNTERS_FILE = "profiles/r2_counters.json"   # written by tools/ncu_counters.py from the ncu captures named inside it


def load_counters():
    """What the rollout kernel EXECUTES per rollout-step and what it moves through DRAM per launch, from ncu captures of the
    shipped kernels (tools/ncu_counters.py -> profiles/r2_counters.json). Nothing here is a literal of this file."""
    try:
        return json.load(open(os.path.join(ROOT, COUNTERS_FILE)))
    except Exception:
        return {}


def assisted_params(energy=True, links=None):
    am = abi.default_assisted_manipulation()
    am.enable_energy_limit = int(energy)
    am.link_position_mode = abi.LINKS_BODY_COM if links is None else links
    return am


def constant_wrench(T, force=(10.0, 0.0, 0.0)):
    w = np.zeros((T, 6))
    w[:, :3] = force
    return w


def workload(name, n_gpus):
    if name == "cfg2":
        return dict(key=name, name="franka_ridgeback_trackpoint_K4096xT64_fp64", system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_TRACK_POINT,
                    params=abi.default_track_point(), K=4096 * n_gpus, horison=0.64, precision=abi.FP64, dtype="f64", x0=abi.huddled_state(), wrench=None)
    if name == "cfg3":
        return dict(key=name, name="franka_ridgeback_assisted_K16384xT128_fp32", system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_ASSISTED_MANIPULATION,
                    params=assisted_params(), K=16384 * n_gpus, horison=1.28, precision=abi.FP32, dtype="f32",
                    x0=abi.huddled_state(10.0), wrench=constant_wrench(128))
    if name == "toy":
        return dict(key=name, name="toy_double_integrator_K1024xT100_fp64", system=abi.SYSTEM_TOY, objective=abi.OBJECTIVE_TOY, params=abi.default_toy_objective(),
                    K=1024 * n_gpus, horison=1.0, precision=abi.FP64, dtype="f64", x0=np.zeros(4), wrench=None)
    if name in ("cfg4_f32", "cfg4_f64"):
        # BASELINE.json config 4: K = 1 048 576 x T = 64 sharded over the ranks (total work fixed: strong scaling)
        f32 = name.endswith("f32")
        return dict(key=name, name="franka_ridgeback_trackpoint_K1048576xT64_" + ("fp32" if f32 else "fp64"), system=abi.SYSTEM_FRANKA_RIDGEBACK,
                    objective=abi.OBJECTIVE_TRACK_POINT, params=abi.default_track_point(), K=1048576, horison=0.64,
                    precision=abi.FP32 if f32 else abi.FP64, dtype="f32" if f32 else "f64", x0=abi.huddled_state(), wrench=None, scaling="strong")
    raise SystemExit("unknown workload " + name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md). One sampler per job (rank 0)
    covering the first `gpus` devices in a single query every `period` s (nvidia-smi takes driver-wide locks; eight ranks
    polling it every 50 ms is load the measurement does not need)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpus=1, enabled=True, period=0.2):
        super().__init__(daemon=True)
        self.gpus, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.enabled, self.period = max(int(gpus), 1), [], set(), False, None, enabled, period

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while self.enabled and not self.stop_flag:
            try:
                ids = ",".join(str(i) for i in range(self.gpus))
                lines = subprocess.run(["nvidia-smi", "-i", ids, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
                clocks = []
                for line in lines:
                    out = line.split(",")
                    clocks.append(float(out[0]))
                    self.max_mhz = float(out[1])
                    for n, v in zip(names, out[2:6]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(n)
                if clocks:
                    self.samples.append(min(clocks))   # the slowest device of the job
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        s = self.samples
        return {"sm_mhz": float(np.median(s)) if s else None, "sm_min_mhz": float(np.min(s)) if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def run_oracle(wl, steps, warmup, threads, budget_s=None):
    """The reference's CPU algorithm (oracle port) on the host cores; returns per-update seconds."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    lib = ol.load()
    holder = abi.make_config(wl["system"], wl["objective"], wl["K"], wl["horison"], keep_best=0, threads=threads)
    o = ol.Oracle(lib, holder, wl["params"])
    times = []
    t_begin = time.perf_counter()
    for u in range(warmup + steps):
        t0 = time.perf_counter()
        rc = o.update(wl["x0"], 0.05 * u, wl["wrench"], None)  # own mt19937 sampling, like the reference
        t1 = time.perf_counter()
        assert rc == 0
        if u >= warmup:
            times.append(t1 - t0)
        if budget_s is not None and u >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    T, R = o.query(abi.QUERY_STEP_COUNT), o.query(abi.QUERY_ROLLOUT_COUNT)
    phases = np.zeros(4)
    lib.oracle_phase_seconds(o.h, ol.ptr(phases))
    o.close()
    return np.array(times), T, R, phases


class Job:
    """What every measurement of this file shares: the ranks of the job, their barrier, the L2 flush buffer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.n_gpus = max(args.gpus, 1)
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.flush = None if args.no_l2_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        self.exchange = args.exchange

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def engine(self, wl):
        """an engine of this workload sharded over the job's ranks (world = 1: the whole rollout set)"""
        from assistedmanipulation_b200 import engine as el
        holder = abi.make_config(wl["system"], wl["objective"], wl["K"], wl["horison"], precision=wl["precision"], dynamics_mode=abi.DYNAMICS_FUSED,
                                 keep_best=0, device=self.local_rank, rank=self.rank, world_size=self.world)
        e = el.Engine(holder, wl["params"])
        if self.world > 1:
            torch, dist = self.torch, self.dist
            rc = abi.connect_ranks(e.lib, e.h, dist, torch, self.exchange)
            ok = torch.tensor([1 if rc == 0 else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok[0]) == 0 and self.exchange == "p2p":
                # peer mappings not available on this box (no IPC / no P2P): every rank falls back to NCCL together
                sys.stderr.write("bench: peer-memory exchange unavailable (%s); using NCCL\n" % e.error())
                e.close()
                e = el.Engine(holder, wl["params"])
                self.exchange = "nccl"
                assert abi.connect_ranks(e.lib, e.h, dist, torch, self.exchange) == 0, e.error()
            else:
                assert int(ok[0]) == 1, e.error()
        return e

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(job, wl, steps, warmup, stage_updates=20, complete_updates=0, clocks=False):
    """`steps` timed updates of one workload on the job's ranks. Device time: CUDA events on the engine's stream around
    each update (inputs resident); wall time: host clock around the C-ABI call (host state in, host control sequence out).
    Returns a dict of raw numbers; the totals are the max over ranks."""
    torch = job.torch
    e = job.engine(wl)
    T, R, nu = e.query(abi.QUERY_STEP_COUNT), e.query(abi.QUERY_ROLLOUT_COUNT), e.query(abi.QUERY_CONTROL_DOF)
    k_local = e.query(abi.QUERY_LOCAL_COUNT)
    # the L2 flush is enqueued on the ENGINE's stream (ordered before the update, outside its CUDA-event
    # window), so no host synchronisation is needed between timed iterations and the ranks stay in step
    sp = C.c_void_p()
    assert e.lib.mppi_b200_stream(e.h, C.byref(sp)) == 0
    engine_stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", job.local_rank))
    x0, wrench = wl["x0"], wl["wrench"]
    step = 0
    for _ in range(warmup):
        assert e.update(x0, 0.05 * step, wrench, seed=1) == 0, e.error()
        step += 1
    launches0 = e.query(abi.QUERY_KERNEL_LAUNCHES)
    # one GPU: every 20 ms; several: every 0.25 s (nvidia-smi takes driver-wide locks — at 8 ranks a 20 ms poll put ~1 ms
    # stalls into one update in a hundred)
    sampler = ClockSampler(job.n_gpus, enabled=(clocks and job.rank == 0), period=0.02 if (job.world == 1 and steps * 3e-4 < 2.0) else 0.25)
    sampler.start()
    gc.collect(); gc.disable()   # no collector pauses of the measuring process inside the timed region
    job.barrier()
    dev_s, wall_s = [], []
    t_region0 = time.perf_counter()
    for _ in range(steps):
        if job.flush is not None:
            with torch.cuda.stream(engine_stream):
                job.flush.fill_(step & 0xff)      # evict L2 between timed iterations
            engine_stream.synchronize()           # keep the flush out of the host-clock (e2e) window too
        if job.world > 1 and job.args.step_barrier:
            job.dist.barrier()                    # optional (see --step-barrier)
        t0 = time.perf_counter()
        rc = e.update(x0, 0.05 * step, wrench, seed=1)   # host state in, host control sequence out
        t1 = time.perf_counter()
        assert rc == 0, e.error()
        wall_s.append(t1 - t0)
        dev_s.append(e.device_seconds())
        step += 1
    job.barrier()
    t_region1 = time.perf_counter()
    gc.enable()
    sampler.stop_flag = True
    sampler.join()
    launches = e.query(abi.QUERY_KERNEL_LAUNCHES) - launches0
    dev_s, wall_s = np.array(dev_s), np.array(wall_s)

    # "complete" latency: the update AND the optimal re-rollout of Trajectory::filter (mppi.cpp:174-176, 450-479), which
    # this engine evaluates on demand — when the optimal cost / breakdown is first read after an update (nothing the caller
    # steers with depends on it; only the logger reads it)
    complete_s = []
    for _ in range(complete_updates):
        if job.flush is not None:
            with torch.cuda.stream(engine_stream):
                job.flush.fill_(step & 0xff)
            engine_stream.synchronize()
        t0 = time.perf_counter()
        assert e.update(x0, 0.05 * step, wrench, seed=1) == 0, e.error()
        e.read(abi.READ_OPTIMAL_COST, 1)                  # Trajectory::get_optimal_total_cost(): runs the re-rollout of this update
        complete_s.append(time.perf_counter() - t0)
        step += 1

    # per-stage device times (separate short loop: the extra events are not in the timed region)
    stage = np.zeros((max(stage_updates, 1), len(abi.STAGES)))
    if stage_updates:
        assert e.lib.mppi_b200_set_profiling(e.h, 1) == 0
        for i in range(stage_updates):
            assert e.update(x0, 0.05 * step, wrench, seed=1) == 0
            step += 1
            e.lib.mppi_b200_stage_seconds(e.h, stage[i].ctypes.data_as(C.POINTER(C.c_double)), len(abi.STAGES))
        e.lib.mppi_b200_set_profiling(e.h, 0)
    stage = np.median(stage, axis=0)
    fma_peak = C.c_double()
    assert e.lib.mppi_b200_measure_fma_peak(job.local_rank, wl["precision"], C.byref(fma_peak)) == 0
    e.close()
    t_dev, t_wall = job.max_over_ranks([dev_s.sum(), wall_s.sum()])
    return dict(T=T, R=R, nu=nu, k_local=k_local, dev_s=dev_s, wall_s=wall_s, t_dev=t_dev, t_wall=t_wall, stage=stage, launches=int(launches),
                clocks=sampler.result() if clocks else None, fma_peak=fma_peak.value, complete_s=np.array(complete_s), wall_region_s=t_region1 - t_region0)


def rooflines(wl, m, peaks, counters):
    """roofline of the rollout kernel (FP pipe) and of the two HBM kernels, from one measurement's stage times"""
    key = wl["key"]
    esz = 8 if wl["precision"] == abi.FP64 else 4
    rollout_s = float(m["stage"][abi.STAGES.index("rollout")])
    algorithmic = FLOPS_PER_STEP.get(key, 9000.0)
    achieved = algorithmic * m["k_local"] * m["T"] / rollout_s / 1e12
    cnt = counters.get(key, {})
    executed = cnt.get("executed_flops_per_rollout_step")
    noise_bytes = m["k_local"] * m["T"] * m["nu"] * esz
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out = {"roofline": {"kernel": "k_rollout", "bound": "fp64_pipe" if wl["precision"] == abi.FP64 else "fp32_pipe", "achieved": achieved, "peak": m["fma_peak"],
                        "unit": "TFLOP/s", "frac": achieved / m["fma_peak"],
                        # DRAM bytes of one launch of this size: the captured bytes per rollout-step x this launch's rollout-steps
                        "traffic": (cnt["rollout_dram_bytes_per_rollout_step"] * m["k_local"] * m["T"]) if "rollout_dram_bytes_per_rollout_step" in cnt else None,
                        "peak_source": "FMA-chain microbenchmark run in this process (mppi_b200_measure_fma_peak); MEASURED_PEAKS.json has no vector FP peak",
                        "algorithmic_flops_per_rollout_step": algorithmic, "executed_flops_per_rollout_step": executed,
                        "executed_frac": (executed * m["k_local"] * m["T"] / rollout_s / 1e12 / m["fma_peak"]) if executed else None,
                        "counters_source": (COUNTERS_FILE + ": " + cnt.get("source", "")) if cnt else None},
           "roofline_hbm": {"peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                            "note": ("the noise buffer (%.1f MB) is produced and consumed inside one update and fits the 126 MB L2" if noise_bytes < 100e6 else
                                     "the noise buffer (%.1f MB) streams through HBM: written once by the sampling kernel, read by the rollout and once more by the weighted sum") % (noise_bytes / 1e6)}}
    for k in ("sample", "weighted_sum"):
        a = noise_bytes / float(m["stage"][abi.STAGES.index(k)]) / 1e9
        out["roofline_hbm"][k] = {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak}
    return out


def secondary_record(job, name, updates, peaks, counters):
    """a few updates of another BASELINE.json configuration: driver-visible evidence, not part of the headline value"""
    wl = workload(name, 1)   # configs 3 and 4 are quoted at fixed size (sharded over however many ranks there are)
    m = measure(job, wl, updates, 3, stage_updates=6)
    if job.rank != 0:
        return None
    units = m["R"] * m["T"] * updates
    rec = {"workload": wl["name"], "n_gpus": job.world, "updates": updates, "ms_per_update": m["t_dev"] / updates * 1e3, "rollout_steps_per_s": units / m["t_dev"],
           "e2e_rollout_steps_per_s": units / m["t_wall"], "scaling": "strong" if job.world > 1 else "single", "dtype": wl["dtype"],
           "rollouts_per_gpu": m["k_local"], "stages_us": {n: float(s * 1e6) for n, s in zip(abi.STAGES, m["stage"])}}
    rec.update(rooflines(wl, m, peaks, counters))
    return rec


def sharded_parity(job):
    """One Philox update of K = 4096 N rollouts sharded over the N ranks against the same rollout set on rank 0 alone:
    per-rollout costs and the best rollout must be bit-equal (the noise depends on the global index only), the published
    control sequence agrees to 1e-9 (the weighted sums are combined in a different order)."""
    from assistedmanipulation_b200 import engine as el
    wl = workload("cfg2", job.world)
    T, nu = 64, 12
    e = job.engine(wl)
    assert e.update(wl["x0"], 0.0, None, seed=77) == 0, e.error()
    mine = e.read(abi.READ_COSTS, e.query(abi.QUERY_LOCAL_COUNT))
    begin = e.query(abi.QUERY_LOCAL_BEGIN)
    U_sharded, argmin_sharded = e.read(abi.READ_OPTIMAL, nu * T), e.query(abi.QUERY_ARGMIN)
    e.close()
    gathered = [None] * job.world
    job.dist.all_gather_object(gathered, (begin, mine, U_sharded, argmin_sharded))
    if job.rank != 0:
        job.barrier()
        return None
    single = el.Engine(abi.make_config(wl["system"], wl["objective"], wl["K"], wl["horison"], precision=wl["precision"], dynamics_mode=abi.DYNAMICS_FUSED,
                                       keep_best=0, device=job.local_rank), wl["params"])
    assert single.update(wl["x0"], 0.0, None, seed=77) == 0, single.error()
    costs = single.read(abi.READ_COSTS, wl["K"] + 2)
    U, argmin = single.read(abi.READ_OPTIMAL, nu * T), single.query(abi.QUERY_ARGMIN)
    single.close()
    costs_equal = all(np.array_equal(c, costs[b:b + len(c)]) for b, c, _, _ in gathered) and sum(len(c) for _, c, _, _ in gathered) == len(costs)
    u_err = max(float(np.abs(u - U).max() / np.abs(U).max()) for _, _, u, _ in gathered)
    ranks_agree = all(np.array_equal(u, gathered[0][2]) for _, _, u, _ in gathered)
    argmin_equal = all(a == argmin for _, _, _, a in gathered)
    job.barrier()
    return {"rollouts": wl["K"], "ranks": job.world, "costs_bit_equal": bool(costs_equal), "argmin_equal": bool(argmin_equal), "control_sequence_rel_err": u_err,
            "ranks_publish_identical_sequences": bool(ranks_agree), "ok": bool(costs_equal and argmin_equal and ranks_agree and u_err <= 1e-9)}


def bench_controllers(job, args, ticks_timed, warmup):
    """BASELINE.json config 5: 256 independent Franka+Ridgeback controllers (K=1024 x T=64 each, assisted-
    manipulation objective, FP32 fast mode, per-controller state and forecast-wrench table), 256/N per GPU,
    no collective. One tick = every controller updates once: all launches first, then all waits, so the
    updates overlap on the device (mppi_b200_update_launch / _wait)."""
    from assistedmanipulation_b200 import engine as el
    rank, world, local_rank = job.rank, job.world, job.local_rank
    total, K, T = 256, 1024, 64
    mine = list(range(rank, total, world))
    params = assisted_params()
    states, wrenches = [], []
    for c in mine:
        x0 = abi.huddled_state(10.0)
        x0[0] += 0.002 * c; x0[1] -= 0.001 * c; x0[2] += 0.003 * c     # per-controller base offset
        states.append(x0)
        ang = 2 * np.pi * c / total + 0.5 * np.arange(T) * 0.01        # a force vector that turns over the horizon
        w = np.zeros((T, 6)); w[:, 0] = 10 * np.cos(ang); w[:, 1] = 10 * np.sin(ang)
        wrenches.append(w)
    states, wrenches = np.ascontiguousarray(np.stack(states)), np.ascontiguousarray(np.stack(wrenches))
    if args.separate_engines:
        # one engine (own streams, own CUDA graph) per controller, overlapped with update_launch / update_wait
        mk = lambda: abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 0.64, precision=abi.FP32,
                                     dynamics_mode=abi.DYNAMICS_FUSED, keep_best=0, device=local_rank)
        engines = [el.Engine(mk(), params) for _ in mine]
    else:
        # ONE batched engine: every kernel runs once with blockIdx.y = controller
        engines = [el.Engine(abi.make_config(abi.SYSTEM_FRANKA_RIDGEBACK, abi.OBJECTIVE_ASSISTED_MANIPULATION, K, 0.64, precision=abi.FP32,
                                             dynamics_mode=abi.DYNAMICS_FUSED, keep_best=0, device=local_rank, batch=len(mine)), params)]
    lib = engines[0].lib
    forecaster = None
    if args.forecast != "table":
        # the wrench table is produced on the device from each controller's MEASURED wrench by the
        # batched forecast producer (Kalman order 1, the simulation's default) and never visits the host
        assert not args.separate_engines, "--forecast kalman needs the batched engine"
        from assistedmanipulation_b200 import forecast as fl
        forecaster = fl.DeviceForecast(abi.FORECAST_KALMAN, 1.0, 0.01, 1, batch=len(mine), device=local_rank)   # on the engine's GPU
        measured = np.ascontiguousarray(wrenches[:, 0, :])
        rng = np.random.default_rng(rank)

    def tick(step):
        if forecaster is not None:
            t = 0.05 * step
            m = measured + rng.normal(0, 0.1, measured.shape)
            assert lib.mppi_b200_forecast_update(forecaster.h, el.ptr(m), t) == 0
            assert lib.mppi_b200_set_wrench_device(engines[0].h, forecaster.table_device(t, 0.01, T)) == 0
            assert engines[0].update(states, t, None, seed=1) == 0, engines[0].error()
            return
        if args.separate_engines:
            for i, e in enumerate(engines):
                rc = lib.mppi_b200_update_launch(e.h, el.ptr(states[i]), 0.05 * step, el.ptr(wrenches[i]), None, abi.NOISE_PHILOX, 1 + i)
                assert rc == 0, e.error()
            for e in engines:
                rc = lib.mppi_b200_update_wait(e.h)
                assert rc == 0, e.error()
        else:
            assert engines[0].update(states, 0.05 * step, wrenches, seed=1) == 0, engines[0].error()

    step = 0
    for _ in range(warmup):
        tick(step); step += 1
    launches0 = sum(e.query(abi.QUERY_KERNEL_LAUNCHES) for e in engines)
    sampler = ClockSampler(job.n_gpus, enabled=(rank == 0))
    sampler.start()
    gc.collect(); gc.disable()
    job.barrier()
    ticks = []
    for _ in range(ticks_timed):
        t0 = time.perf_counter()
        tick(step); step += 1
        ticks.append(time.perf_counter() - t0)
    job.barrier()
    gc.enable()
    sampler.stop_flag = True
    sampler.join()
    launches = sum(e.query(abi.QUERY_KERNEL_LAUNCHES) for e in engines) - launches0
    dev = np.array([e.device_seconds() for e in engines])
    (t_wall,) = job.max_over_ranks([float(np.sum(ticks))])
    for e in engines:
        e.close()
    if forecaster is not None:
        forecaster.close()
    if rank != 0:
        return None
    units = total * (K + 2) * T * ticks_timed
    ticks = np.array(ticks)
    return {"metric": "rollout-steps/s", "value": units / t_wall, "unit": "rollout-steps/s", "n_gpus": job.n_gpus, "steps": ticks_timed, "warmup": warmup,
            "ms_per_step": t_wall / ticks_timed * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "256_controllers_franka_ridgeback_assisted_K1024xT64_fp32", "controllers": total, "controllers_per_gpu": len(mine),
                       "rollouts": K, "steps_per_rollout": T, "noise": "in-kernel Philox4x32-10", "parallelism": "controllers split over %d GPU(s), no collective; %s" % (world, "one engine per controller, overlapped streams" if args.separate_engines else "one batched engine per GPU (blockIdx.y = controller)"),
                       "timing": "host clock around one tick (launch all controllers, wait for all); per-controller states and wrench tables come from host memory every tick" if forecaster is None else
                       "host clock around one tick (measured wrenches H2D -> batched Kalman forecast update -> device wrench table -> batched MPPI update)",
                       "forecast": "host tables" if forecaster is None else "device Kalman producer, order 1, dt 0.01, horison 1.0"},
            "clocks": sampler.result(),
            "e2e": {"value": units / t_wall, "unit": "rollout-steps/s", "h2d_bytes_per_step": int(len(mine) * 8 * (40 + (6 * T if forecaster is None else 6))), "d2h_bytes_per_step": int(len(mine) * 8 * (12 * T + 5)),
                    "tick_latency_us": {"p50": float(np.median(ticks) * 1e6), "p99": float(np.percentile(ticks, 99) * 1e6)}},
            "gpu_launches": int(launches) + (2 * ticks_timed if forecaster is not None else 0),   # + k_kalman_update, k_table per tick
            "per_controller_device_update_us": {"p50": float(np.median(dev) * 1e6), "max": float(dev.max() * 1e6)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the block of short runs of the other BASELINE.json configurations")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--forecast", default="table", choices=["table", "kalman"], help="cfg5: host wrench tables, or the device forecast producer fed measured wrenches")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: the library's own kernels over NVLink peer memory, or NCCL all-reduces")
    ap.add_argument("--step-barrier", action="store_true",
                    help="N > 1: a torch.distributed barrier before every timed update. Default off: the ranks are coupled by the update's own exchanges (a rank "
                         "cannot run ahead of its peers), and at 8 ranks the per-step NCCL barrier itself put ~1 ms stalls into 1 % of the updates "
                         "(measured: p99 1266 us with it, 238 us without, p50 226 / 223 us)")
    ap.add_argument("--separate-engines", action="store_true", help="cfg5: one engine per controller instead of one batched engine")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    n_gpus = max(args.gpus, 1)
    args.warmup = max(args.warmup, 3)
    wl = workload(args.workload if args.workload != "cfg5" else "cfg3", n_gpus)
    host_threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample per step: one full update of the same workload at N=1 size per GPU
        times, T, R, phases = run_oracle(wl, args.steps, min(args.warmup, 3), host_threads, budget_s=150.0)   # bounded: stops after ~150 s of host work
        value = R * T / times.mean()
        line = {"impl": "reference", "metric": "rollout-steps/s", "value": value, "unit": "rollout-steps/s", "n_gpus": n_gpus, "steps": len(times),
                "warmup": args.warmup, "ms_per_step": times.mean() * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                # the same keys as the GPU arm's config (the driver compares the two dicts)
                "config": {"workload": wl["name"], "rollouts": wl["K"], "static_rollouts": 2, "steps_per_rollout": T, "time_step": 0.01, "update_cadence_s": 0.05,
                           "noise": "mt19937 gaussian (controller/gaussian.hpp)", "dynamics_mode": "faithful (RNEA + ABA like pinocchio_dynamics.cpp:153-224)", "keep_best_rollouts": 0,
                           "parallelism": "reference thread pool partition over %d host threads (mppi.cpp:272-307)" % host_threads,
                           "l2": "n/a (host)", "timing": "host clock around Trajectory::update of the oracle port"},
                "latency_us": {"p50": float(np.median(times) * 1e6), "p99": float(np.percentile(times, 99) * 1e6)},
                "cpu_baseline": {"value": value, "unit": "rollout-steps/s", "cores": host_threads, "kind": "port",
                                 "sample": "%d full updates of the workload (K+2=%d rollouts x T=%d), Trajectory::filter included" % (len(times), R, T),
                                 "phase_split_s": dict(zip(["sample", "rollout", "optimise", "filter"], [float(x) for x in phases]))},
                "e2e": {"value": value, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    job = Job(args)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    counters = load_counters()
    if args.workload == "cfg5":
        line = bench_controllers(job, args, args.steps, args.warmup)
        job.close()
        if line is not None:
            print(json.dumps(line))
        return 0

    m = measure(job, wl, args.steps, args.warmup, stage_updates=20, complete_updates=50 if job.world == 1 else 0, clocks=True)
    line = None
    if job.rank == 0:
        T, R, nu = m["T"], m["R"], m["nu"]
        units = R * T * args.steps
        dev_s, wall_s = m["dev_s"], m["wall_s"]
        line = {
            "metric": "rollout-steps/s", "value": units / m["t_dev"], "unit": "rollout-steps/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["t_dev"] / args.steps * 1e3, "higher_is_better": True, "scaling": wl.get("scaling", "weak"), "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["name"], "rollouts": wl["K"], "static_rollouts": 2, "steps_per_rollout": T, "time_step": 0.01, "update_cadence_s": 0.05,
                       "noise": "in-kernel Philox4x32-10", "dynamics_mode": "fused", "keep_best_rollouts": 0,
                       "parallelism": "rollouts sharded over %d GPU(s); %s of [-min,max] and [sum w, sum w*eps]" % (job.world, "no exchange" if job.world == 1 else ("the library's own kernels over NVLink peer memory (IPC mailboxes)" if job.exchange == "p2p" else "NCCL all-reduce")),
                       "l2": "not flushed" if job.flush is None else "flushed between timed iterations (256 MiB write)",
                       "timing": "value: CUDA events on the engine stream around each update (inputs resident); e2e: host clock around the C-ABI call"},
            "clocks": m["clocks"],
            "e2e": {"value": units / m["t_wall"], "unit": "rollout-steps/s", "h2d_bytes_per_step": int(8 * (40 + 6 * T)), "d2h_bytes_per_step": int(8 * (nu * T + 4)),
                    "update_latency_us": {"p50": float(np.median(wall_s) * 1e6), "p99": float(np.percentile(wall_s, 99) * 1e6), "max": float(wall_s.max() * 1e6),
                                          "samples": int(len(wall_s)), "p99_note": "percentile over `samples` consecutive updates; with fewer than ~200 samples it is close to the maximum"}},
            "gpu_launches": m["launches"],
            "device_update_us": {"p50": float(np.median(dev_s) * 1e6), "p99": float(np.percentile(dev_s, 99) * 1e6)},
            "stages_us": {n: float(s * 1e6) for n, s in zip(abi.STAGES, m["stage"])},
            "wall_region_s": m["wall_region_s"],
        }
        if len(m["complete_s"]):
            c = m["complete_s"]
            line["e2e_complete"] = {"update_latency_us": {"p50": float(np.median(c) * 1e6), "p99": float(np.percentile(c, 99) * 1e6), "samples": int(len(c))},
                                    "note": "update + get_optimal_total_cost(): the optimal re-rollout of Trajectory::filter (mppi.cpp:174-176) is evaluated on demand, when its result is first read; "
                                            "the published control sequence does not depend on it (only the logger reads it), so `value` and `e2e` do not contain it; the reference arm's update does"}
        line.update(rooflines(wl, m, peaks, counters))

    if not args.no_secondary and args.workload == "cfg2":
        sec = {}
        if job.world == 1:
            for name, updates in (("cfg3", 10), ("cfg4_f32", 10), ("cfg4_f64", 10)):
                sec[name] = secondary_record(job, name, updates, peaks, counters)
        else:
            sec["cfg4_f64"] = secondary_record(job, "cfg4_f64", 5, peaks, counters)
            sec["cfg4_f32"] = secondary_record(job, "cfg4_f32", 5, peaks, counters)
            sec["sharded_parity"] = sharded_parity(job)
        c5 = bench_controllers(job, args, 5, 3)
        if c5 is not None:
            sec["cfg5"] = {k: c5[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "config", "e2e", "per_controller_device_update_us", "gpu_launches")}
        if line is not None:
            line["secondary"] = sec
            line["secondary_note"] = "short runs of the other BASELINE.json configurations in the same process; none of them enters `value`"
    job.close()
    if job.rank != 0:
        return 0
    if n_gpus == 1 and not args.no_cpu_baseline:
        wl1 = workload(args.workload, 1)
        times, T1, R1, phases = run_oracle(wl1, 8, 1, host_threads, budget_s=25.0)
        line["cpu_baseline"] = {"value": R1 * T1 / times.mean(), "unit": "rollout-steps/s", "cores": host_threads, "kind": "port",
                                "sample": "%d full updates of the same workload (K+2=%d x T=%d) on the oracle port, %d pool threads" % (len(times), R1, T1, host_threads),
                                "update_ms": {"p50": float(np.median(times) * 1e3), "max": float(times.max() * 1e3)},
                                "phase_split_s": dict(zip(["sample", "rollout", "optimise", "filter"], [float(x) for x in phases]))}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
