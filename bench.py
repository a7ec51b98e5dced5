#!/usr/bin/env python3
"""bench.py — the MPPI rollout path on N B200s (driver contract, tier ④).

A "step" is one mppi::Trajectory::update() (reference src/controller/mppi.cpp:154-187): sample ->
K+2 rollouts of T steps -> exp-weighting -> weighted-sum update -> smoothing/clamp. Workload at N=1 is
BASELINE.json configs[1]: Franka Research 3 + Ridgeback, TrackPoint objective, K=4096 x T=64, FP64.
For N>1 the rollout set grows with N (K = 4096*N, "weak") and is sharded over the ranks; the two
exchanges of the path (min/max of the costs, weighted sums) are NCCL all-reduces inside the library.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|toy]

`--impl reference` times the reference's CPU algorithm (the oracle port with its thread pool on all
host cores; the reference itself cannot run K > 253 nor be built offline) on the same
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
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from assistedmanipulation_b200 import abi  # noqa: E402

# Algorithmic work per rollout-step: Featherstone operation counts at n = 12 one-dof
# joints — ABA 4641 + RNEA 1880 + second-order FK / frames / WORLD jacobian ~2300 + cost + Euler + tank.
FLOPS_PER_STEP = {"cfg2": 9000.0, "cfg3": 9500.0, "toy": 25.0, "cfg4_f32": 9000.0, "cfg4_f64": 9000.0}
# What the rollout kernel EXECUTES per rollout-step (2*DFMA + DMUL + DADD per thread along the step loop, counted in the
# SASS by tools/sass_cycles.py, profiles/r1_sass_model.txt: 858 / 431 / 212 for the unrolled build, which configs 2 and 4
# both launch (k_rollout.cuh); the loop-body build executes 1287 / 559 / 249 and the round's first kernel executed
# 1197 / 837 / 534 by the ncu source page of profiles/r1_cfg2_final.ncu-rep): the structure-exploiting solver needs far
# fewer operations than the reference's generic algorithm counted above, so the algorithmic fraction can pass 1 where
# the pipe is full.
EXECUTED_FLOPS_PER_STEP = {"cfg2": 2359.0, "cfg4_f64": 2359.0}
# dram__bytes_read.sum + dram__bytes_write.sum of one k_rollout launch, same capture (the noise rows, read once)
ROLLOUT_DRAM_BYTES = {"cfg2": 25.23e6}


def workload(name, n_gpus):
    if name == "cfg2":
        return dict(name="franka_ridgeback_trackpoint_K4096xT64_fp64", system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_TRACK_POINT,
                    params=abi.default_track_point(), K=4096 * n_gpus, horison=0.64, precision=abi.FP64, dtype="f64", x0=abi.huddled_state(), wrench=None)
    if name == "cfg3":
        import cases
        return dict(name="franka_ridgeback_assisted_K16384xT128_fp32", system=abi.SYSTEM_FRANKA_RIDGEBACK, objective=abi.OBJECTIVE_ASSISTED_MANIPULATION,
                    params=cases.assisted_params(True, abi.LINKS_BODY_COM), K=16384 * n_gpus, horison=1.28, precision=abi.FP32, dtype="f32",
                    x0=abi.huddled_state(10.0), wrench=cases.constant_wrench(128))
    if name == "toy":
        return dict(name="toy_double_integrator_K1024xT100_fp64", system=abi.SYSTEM_TOY, objective=abi.OBJECTIVE_TOY, params=abi.default_toy_objective(),
                    K=1024 * n_gpus, horison=1.0, precision=abi.FP64, dtype="f64", x0=np.zeros(4), wrench=None)
    if name in ("cfg4_f32", "cfg4_f64"):
        # BASELINE.json config 4: K = 1 048 576 x T = 64 sharded over the ranks (total work fixed: strong scaling)
        f32 = name.endswith("f32")
        return dict(name="franka_ridgeback_trackpoint_K1048576xT64_" + ("fp32" if f32 else "fp64"), system=abi.SYSTEM_FRANKA_RIDGEBACK,
                    objective=abi.OBJECTIVE_TRACK_POINT, params=abi.default_track_point(), K=1048576, horison=0.64,
                    precision=abi.FP32 if f32 else abi.FP64, dtype="f32" if f32 else "f64", x0=abi.huddled_state(), wrench=None, scaling="strong")
    raise SystemExit("unknown workload " + name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md). One sampler per job (rank 0)
    covering the first `gpus` devices in a single query every 0.2 s (nvidia-smi takes driver-wide locks; eight ranks polling
    it every 50 ms is load the measurement does not need)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpus=1, enabled=True):
        super().__init__(daemon=True)
        self.gpus, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.enabled = max(int(gpus), 1), [], set(), False, None, enabled

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
            time.sleep(0.2)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_oracle(wl, steps, warmup, threads, budget_s=None):
    """The reference's CPU algorithm (oracle port) on the host cores; returns per-update seconds."""
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


def bench_controllers(args, rank, world, local_rank, n_gpus, dist, torch, el):
    """BASELINE.json config 5: 256 independent Franka+Ridgeback controllers (K=1024 x T=64 each, assisted-
    manipulation objective, FP32 fast mode, per-controller state and forecast-wrench table), 256/N per GPU,
    no collective. One tick = every controller updates once: all launches first, then all waits, so the
    updates overlap on the device (mppi_b200_update_launch / _wait)."""
    import cases
    total, K, T = 256, 1024, 64
    mine = list(range(rank, total, world))
    params = cases.assisted_params(True, abi.LINKS_BODY_COM)
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
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import forecast_lib as fl
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = 0
    for _ in range(args.warmup):
        tick(step); step += 1
    launches0 = sum(e.query(abi.QUERY_KERNEL_LAUNCHES) for e in engines)
    sampler = ClockSampler(n_gpus, enabled=(rank == 0))
    sampler.start()
    gc.collect(); gc.disable()   # no collector pauses of the measuring process inside the timed region
    barrier()
    ticks = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        tick(step); step += 1
        ticks.append(time.perf_counter() - t0)
    barrier()
    gc.enable()
    sampler.stop_flag = True
    sampler.join()
    launches = sum(e.query(abi.QUERY_KERNEL_LAUNCHES) for e in engines) - launches0
    dev = np.array([e.device_seconds() for e in engines])
    t_wall = float(np.sum(ticks))
    if world > 1:
        tt = torch.tensor([t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_wall = float(tt[0])
        dist.destroy_process_group()
    for e in engines:
        e.close()
    if forecaster is not None:
        forecaster.close()
    if rank != 0:
        return 0
    units = total * (K + 2) * T * args.steps
    ticks = np.array(ticks)
    line = {"metric": "rollout-steps/s", "value": units / t_wall, "unit": "rollout-steps/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "256_controllers_franka_ridgeback_assisted_K1024xT64_fp32", "controllers": total, "controllers_per_gpu": len(mine),
                       "rollouts": K, "steps_per_rollout": T, "noise": "in-kernel Philox4x32-10", "parallelism": "controllers split over %d GPU(s), no collective; %s" % (world, "one engine per controller, overlapped streams" if args.separate_engines else "one batched engine per GPU (blockIdx.y = controller)"),
                       "timing": "host clock around one tick (launch all controllers, wait for all); per-controller states and wrench tables come from host memory every tick" if forecaster is None else
                       "host clock around one tick (measured wrenches H2D -> batched Kalman forecast update -> device wrench table -> batched MPPI update)",
                       "forecast": "host tables" if forecaster is None else "device Kalman producer, order 1, dt 0.01, horison 1.0"},
            "clocks": sampler.result(),
            "e2e": {"value": units / t_wall, "unit": "rollout-steps/s", "h2d_bytes_per_step": int(len(mine) * 8 * (40 + (6 * T if forecaster is None else 6))), "d2h_bytes_per_step": int(len(mine) * 8 * (12 * T + 5)),
                    "tick_latency_us": {"p50": float(np.median(ticks) * 1e6), "p99": float(np.percentile(ticks, 99) * 1e6)}},
            "gpu_launches": int(launches) + (2 * args.steps if forecaster is not None else 0),   # + k_kalman_update, k_table per tick
            "per_controller_device_update_us": {"p50": float(np.median(dev) * 1e6), "max": float(dev.max() * 1e6)}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--forecast", default="table", choices=["table", "kalman"], help="cfg5: host wrench tables, or the device forecast producer fed measured wrenches")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: the library's own kernels over NVLink peer memory, or NCCL all-reduces")
    ap.add_argument("--separate-engines", action="store_true", help="cfg5: one engine per controller instead of one batched engine")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
                "config": {"workload": wl["name"], "rollouts": wl["K"], "steps_per_rollout": T, "noise": "mt19937 gaussian (controller/gaussian.hpp)"},
                "latency_us": {"p50": float(np.median(times) * 1e6), "p99": float(np.percentile(times, 99) * 1e6)},
                "cpu_baseline": {"value": value, "unit": "rollout-steps/s", "cores": host_threads, "kind": "port",
                                 "sample": "%d full updates of the workload (K+2=%d rollouts x T=%d)" % (len(times), R, T),
                                 "phase_split_s": dict(zip(["sample", "rollout", "optimise", "filter"], [float(x) for x in phases]))},
                "e2e": {"value": value, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import engine_lib as el
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload == "cfg5":
        return bench_controllers(args, rank, world, local_rank, n_gpus, dist, torch, el)
    holder = abi.make_config(wl["system"], wl["objective"], wl["K"], wl["horison"], precision=wl["precision"], dynamics_mode=abi.DYNAMICS_FUSED,
                             keep_best=0, device=local_rank, rank=rank, world_size=world)
    e = el.Engine(holder, wl["params"])
    exchange = args.exchange
    if world > 1:
        rc = abi.connect_ranks(e.lib, e.h, dist, torch, exchange)
        ok = torch.tensor([1 if rc == 0 else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) == 0 and exchange == "p2p":
            # peer mappings not available on this box (no IPC / no P2P): every rank falls back to NCCL together
            sys.stderr.write("bench: peer-memory exchange unavailable (%s); using NCCL\n" % e.error())
            e.close()
            e = el.Engine(holder, wl["params"])
            exchange = "nccl"
            assert abi.connect_ranks(e.lib, e.h, dist, torch, exchange) == 0, e.error()
        else:
            assert int(ok[0]) == 1, e.error()
    T, R = e.query(abi.QUERY_STEP_COUNT), e.query(abi.QUERY_ROLLOUT_COUNT)
    nu = e.query(abi.QUERY_CONTROL_DOF)
    flush = None if args.no_l2_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    # the L2 flush is enqueued on the ENGINE's stream (ordered before the update, outside its CUDA-event
    # window), so no host synchronisation is needed between timed iterations and the ranks stay in step
    sp = C.c_void_p()
    assert e.lib.mppi_b200_stream(e.h, C.byref(sp)) == 0
    engine_stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    x0, wrench = wl["x0"], wl["wrench"]
    step = 0
    for _ in range(args.warmup):
        assert e.update(x0, 0.05 * step, wrench, seed=1) == 0, e.error()
        step += 1
    launches0 = e.query(abi.QUERY_KERNEL_LAUNCHES)
    sampler = ClockSampler(n_gpus, enabled=(rank == 0))
    sampler.start()
    gc.collect(); gc.disable()   # no collector pauses of the measuring process inside the timed region
    barrier()
    dev_s, wall_s = [], []
    t_region0 = time.perf_counter()
    for _ in range(args.steps):
        if flush is not None:
            with torch.cuda.stream(engine_stream):
                flush.fill_(step & 0xff)      # evict L2 between timed iterations
            engine_stream.synchronize()       # keep the flush out of the host-clock (e2e) window too
        if world > 1:
            dist.barrier()                    # ranks enter the step together: the in-step all-reduces then measure exchange cost, not host skew
        t0 = time.perf_counter()
        rc = e.update(x0, 0.05 * step, wrench, seed=1)   # host state in, host control sequence out
        t1 = time.perf_counter()
        assert rc == 0, e.error()
        wall_s.append(t1 - t0)
        dev_s.append(e.device_seconds())
        step += 1
    barrier()
    t_region1 = time.perf_counter()
    gc.enable()
    sampler.stop_flag = True
    sampler.join()
    launches = e.query(abi.QUERY_KERNEL_LAUNCHES) - launches0
    dev_s, wall_s = np.array(dev_s), np.array(wall_s)

    # per-stage device times (separate short loop: the extra events are not in the timed region)
    assert e.lib.mppi_b200_set_profiling(e.h, 1) == 0
    stage = np.zeros((20, len(abi.STAGES)))
    for i in range(20):
        assert e.update(x0, 0.05 * step, wrench, seed=1) == 0
        step += 1
        e.lib.mppi_b200_stage_seconds(e.h, stage[i].ctypes.data_as(C.POINTER(C.c_double)), len(abi.STAGES))
    stage = np.median(stage, axis=0)
    e.lib.mppi_b200_set_profiling(e.h, 0)

    t_dev, t_wall = float(dev_s.sum()), float(wall_s.sum())
    if world > 1:
        tt = torch.tensor([t_dev, t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_wall = float(tt[0]), float(tt[1])
    if rank != 0:
        e.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    units = R * T * args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fma_peak = C.c_double()
    assert e.lib.mppi_b200_measure_fma_peak(local_rank, wl["precision"], C.byref(fma_peak)) == 0
    rollout_s = float(stage[abi.STAGES.index("rollout")])
    k_local = e.query(abi.QUERY_LOCAL_COUNT)
    achieved_tflops = FLOPS_PER_STEP.get(args.workload, 9000.0) * k_local * T / rollout_s / 1e12
    esz = 8 if wl["precision"] == abi.FP64 else 4
    noise_bytes = k_local * T * nu * esz
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": "rollout-steps/s", "value": units / t_dev, "unit": "rollout-steps/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": wl.get("scaling", "weak"), "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": wl["name"], "rollouts": wl["K"], "static_rollouts": 2, "steps_per_rollout": T, "time_step": 0.01, "update_cadence_s": 0.05,
                   "noise": "in-kernel Philox4x32-10", "dynamics_mode": "fused", "keep_best_rollouts": 0,
                   "parallelism": "rollouts sharded over %d GPU(s); %s of [-min,max] and [sum w, sum w*eps]" % (world, "no exchange" if world == 1 else ("exchange kernels over NVLink peer memory (IPC mailboxes)" if exchange == "p2p" else "NCCL all-reduce")),
                   "l2": "not flushed" if flush is None else "flushed between timed iterations (256 MiB write)",
                   "timing": "value: CUDA events on the engine stream around each update (inputs resident); e2e: host clock around the C-ABI call"},
        "clocks": sampler.result(),
        "e2e": {"value": units / t_wall, "unit": "rollout-steps/s", "h2d_bytes_per_step": int(8 * (40 + 6 * T)), "d2h_bytes_per_step": int(8 * (nu * T + 4)),
                "update_latency_us": {"p50": float(np.median(wall_s) * 1e6), "p99": float(np.percentile(wall_s, 99) * 1e6), "max": float(wall_s.max() * 1e6)}},
        "gpu_launches": int(launches),
        "device_update_us": {"p50": float(np.median(dev_s) * 1e6), "p99": float(np.percentile(dev_s, 99) * 1e6)},
        "stages_us": {n: float(s * 1e6) for n, s in zip(abi.STAGES, stage)},
        "roofline": {"kernel": "k_rollout", "bound": "fp64_pipe" if wl["precision"] == abi.FP64 else "fp32_pipe", "achieved": achieved_tflops, "peak": fma_peak.value,
                     "unit": "TFLOP/s", "frac": achieved_tflops / fma_peak.value, "traffic": ROLLOUT_DRAM_BYTES.get(args.workload),
                     "peak_source": "FMA-chain microbenchmark run in this process (mppi_b200_measure_fma_peak); MEASURED_PEAKS.json has no vector FP peak",
                     "algorithmic_flops_per_rollout_step": FLOPS_PER_STEP.get(args.workload, 9000.0),
                     "executed_flops_per_rollout_step": EXECUTED_FLOPS_PER_STEP.get(args.workload),
                     "executed_frac": (EXECUTED_FLOPS_PER_STEP[args.workload] * k_local * T / rollout_s / 1e12 / fma_peak.value) if args.workload in EXECUTED_FLOPS_PER_STEP else None},
        "roofline_hbm": {
            "sample": {"bound": "hbm", "achieved": noise_bytes / float(stage[abi.STAGES.index("sample")]) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
            "weighted_sum": {"bound": "hbm", "achieved": noise_bytes / float(stage[abi.STAGES.index("weighted_sum")]) / 1e9, "peak": hbm_peak, "unit": "GB/s"},
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "note": ("the noise buffer (%.1f MB) is produced and consumed inside one update and fits the 126 MB L2" if noise_bytes < 100e6 else "the noise buffer (%.1f MB) streams through HBM: written once by the sampling kernel, read by the rollout and once more by the weighted sum") % (noise_bytes / 1e6)},
        "wall_region_s": t_region1 - t_region0,
    }
    for k in ("sample", "weighted_sum"):
        line["roofline_hbm"][k]["frac"] = line["roofline_hbm"][k]["achieved"] / hbm_peak
    e.close()
    if world > 1:
        dist.destroy_process_group()
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
