#!/usr/bin/env python3
"""profiles/r2_counters.json from ncu metric passes of the shipped kernels (no literals in bench.py): for every workload the
FP operations the rollout kernel EXECUTES per rollout-step (SASS thread-instruction counters: 2 per FMA, 1 per MUL / ADD),
its DRAM traffic per launch, and the duration / traffic of the sampling and weighted-sum kernels.

  ncu --csv --metrics <METRICS> --clock-control none --launch-skip S --launch-count C python tools/prof_target.py cfg2 6 > x.csv
  python tools/ncu_counters.py cfg2=gpurun_out/r2_counters_cfg2.csv cfg3=... cfg4_f64=... > profiles/r2_counters.json
"""
import csv
import json
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,"
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum")
# rollouts x steps of the prof_target.py workloads (K + 2 rollouts)
UNITS = {"cfg2": 4098 * 64, "cfg3": 16386 * 128, "cfg4_f64": 131074 * 64, "cfg4_f32": 131074 * 64, "cfg2_f32": 4098 * 64}


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def to_us(value, unit):
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def read(path):
    rows = [r for r in csv.reader(open(path)) if r]
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    iK, iM, iU, iV, iG, iB, iI = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size", "Block Size", "ID"))
    launches = {}
    for r in rows[start + 1:]:
        if len(r) <= iV:
            continue
        L = launches.setdefault(r[iI], {"kernel": r[iK], "grid": r[iG], "block": r[iB]})
        L[r[iM]] = (r[iV], r[iU])
    return list(launches.values())


def summarise(key, path):
    launches = read(path)
    out = {"source": "ncu metric pass " + path.split("/")[-1] + " (tools/prof_target.py " + key + "; kernels as committed)", "kernels": {}}
    for L in launches:
        name = L["kernel"].split("(")[0].replace("void ", "")
        if "k_rollout" in name and L["grid"].replace(" ", "").startswith("(1,"):
            continue   # the one-thread optimal re-rollout (on demand) is not the rollout grid
        flops = 0.0
        for m, w in (("dfma", 2), ("dmul", 1), ("dadd", 1), ("ffma", 2), ("fmul", 1), ("fadd", 1)):
            v = L.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % m)
            if v:
                flops += w * float(v[0].replace(",", ""))
        rec = {"grid": L["grid"], "block": L["block"], "duration_us": to_us(*L["gpu__time_duration.sum"]),
               "dram_bytes": to_bytes(*L["dram__bytes_read.sum"]) + to_bytes(*L["dram__bytes_write.sum"]),
               "warp_instructions": float(L["smsp__inst_executed.sum"][0].replace(",", "")), "fp_operations": flops}
        out["kernels"].setdefault(name, rec)   # first launch of each kernel in the window
        if "k_rollout" in name and "executed_flops_per_rollout_step" not in out:
            out["executed_flops_per_rollout_step"] = flops / UNITS[key]
            out["rollout_dram_bytes_per_launch"] = rec["dram_bytes"]
            out["rollout_dram_bytes_per_rollout_step"] = rec["dram_bytes"] / UNITS[key]
            out["rollout_kernel"] = name
            out["rollout_duration_us_under_ncu"] = rec["duration_us"]
    return out


if __name__ == "__main__":
    result = {}
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        result[key] = summarise(key, path)
    json.dump(result, sys.stdout, indent=1)
    print()
