"""Static issue-cycle model of one kernel from its SASS control words (no GPU needed).

    python tools/sass_cycles.py <object-or-cubin> <substring of the mangled kernel name> [--trip N] [--dump]

Every sm_100a instruction carries, in bits 105..108 of its 128-bit encoding, the number of cycles the scheduler
waits before it issues the next instruction of the same warp; ptxas sets it from the fixed latencies of the
dependency chains (a dependent DFMA pair shows up as a stall of 8 or so). With ONE resident warp per SM sub-partition
(the K = 4096 rollout kernel: 129 warps on 148 SMs) nothing else hides those cycles, so the sum of the stall fields
along the executed path is the kernel's time per step, up to the variable-latency waits (local / global loads,
MUFU) that are tracked by scoreboards instead. The tool finds the loops (backward branches), prints per-loop sums
and a weighted total: the outermost loop of the biggest nest is taken as the step loop and every loop nested in it
is given `--trip` iterations (7: the arm joints). Forward branches are assumed not taken (they guard slow paths).
Calibration against ncu on a B200 is recorded in DESIGN.md section 5."""
import argparse
import re
import subprocess
import sys
import tempfile
import os

INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
HI = re.compile(r"^\s*/\* 0x([0-9a-f]{16}) \*/")


def disassemble(path, pattern, inline=False):
    tmp = None
    if not path.endswith(".cubin"):
        tmp = tempfile.mkdtemp()
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
        assert cubins, "no cubin in " + path
        path = cubins[0]
    text = subprocess.run(["nvdisasm", "-hex", "-g"] + (["-gi"] if inline else []) + [path], check=True, capture_output=True, text=True).stdout.splitlines()
    funcs, cur, name = {}, None, None
    for line in text:
        m = re.match(r"^\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            continue
        if line.startswith("\t.section") or line.startswith("\t.sectio"):
            cur = None
            continue
        if cur is not None:
            cur.append(line)
    hits = [n for n in funcs if pattern in n]
    assert len(hits) == 1, "kernel name pattern matches %d functions: %s" % (len(hits), hits[:8])
    return hits[0], funcs[hits[0]]


def parse(lines):
    """-> list of dicts(addr, text, stall, yield_, wbar, rbar, wait, label) in program order"""
    out, labels, pending = [], {}, None
    where = ("?", 0)
    chain, in_chain = [], False   # with -gi: innermost frame first, every "inlined at" caller after it
    i = 0
    while i < len(lines):
        line = lines[i]
        mf = re.match(r'^\s*//## File "([^"]+)", line (\d+)', line)
        if mf:
            if not in_chain:
                chain = []
            in_chain = True
            frame = (os.path.basename(mf.group(1)), int(mf.group(2)))
            if not chain or chain[-1] != frame:
                chain.append(frame)
            where = chain[0]
            i += 1
            continue
        in_chain = False
        ml = re.match(r"^(\.L_x_\d+):", line)
        if ml:
            pending = ml.group(1)
        m = INS.match(line)
        if m:
            hi = int(HI.match(lines[i + 1]).group(1), 16)
            ins = dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(hi >> 41) & 0xF, yield_=(hi >> 45) & 1,
                       wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F, where=where, chain=tuple(chain))
            if pending:
                labels[pending] = len(out)
                pending = None
            out.append(ins)
            i += 2
            continue
        i += 1
    return out, labels


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0] if t else ""


def loops(ins, labels):
    """backward branches -> (head index, branch index)"""
    found = []
    for k, it in enumerate(ins):
        if opcode(it["text"]).startswith("BRA"):
            m = re.search(r"`\((\.L_x_\d+)\)", it["text"])
            if m and m.group(1) in labels and labels[m.group(1)] <= k:
                found.append((labels[m.group(1)], k))
    return found


FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def is_fp64(it):
    return not it.get("skipped") and opcode(it["text"]).split(".")[0] in FP64_OPS


def cost(it):
    return 0 if it.get("skipped") else max(1, it["stall"])


def summarize(ins, lo, hi):
    seg = ins[lo:hi + 1]
    ops = {}
    for it in seg:
        o = opcode(it["text"]).split(".")[0]
        ops[o] = ops.get(o, 0) + 1
    return dict(n=len(seg), cycles=sum(cost(it) for it in seg), waits=sum(1 for it in seg if it["wait"]), ops=ops)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("pattern")
    ap.add_argument("--trip", type=int, default=7)
    ap.add_argument("--fp64-issue", type=int, default=2, help="cycles one FP64 warp instruction occupies the pipe")
    ap.add_argument("--dump", action="store_true", help="print every instruction of the step loop with its stall field")
    ap.add_argument("--skip", default="", help="comma-separated file:line whose instructions are left out (slow paths behind forward branches)")
    ap.add_argument("--lines", type=int, default=0, help="print the N source lines with the most stall cycles per step (needs -lineinfo)")
    ap.add_argument("--phases", default="", help="source file (e.g. rollout_core.cuh): attribute the step to the calls made from that file, through the inline chains (nvdisasm -gi)")
    ap.add_argument("--phase-depth", type=int, default=2, help="how many frames below the call in --phases are listed")
    a = ap.parse_args()
    name, lines = disassemble(a.path, a.pattern, inline=bool(a.phases))
    ins, labels = parse(lines)
    skip = set(tuple([w.split(":")[0], int(w.split(":")[1])]) for w in a.skip.split(",") if w)
    for it in ins:
        if it["where"] in skip:
            it["stall"], it["skipped"] = 0, True
    lp = loops(ins, labels)
    print("kernel %s: %d instructions, %d loops" % (name, len(ins), len(lp)))
    if not lp:
        s = summarize(ins, 0, len(ins) - 1)
        print("straight line: %d instructions, %d stall cycles" % (s["n"], s["cycles"]))
        return
    outer = max(lp, key=lambda r: r[1] - r[0])
    inner = [r for r in lp if r != outer and outer[0] <= r[0] and r[1] <= outer[1]]
    # nesting depth of every instruction inside the step loop
    total_c = total_n = 0.0
    fp64 = 0.0
    by_line = {}
    mix = {}
    # issue-time walk: an instruction issues `stall` cycles after its predecessor, and an FP64 instruction not before
    # the FP64 pipe of the sub-partition is free again (one warp instruction per 2 cycles, tools/microbench/dfma_issue.cu);
    # the second condition is enforced by the hardware, not by the stall field (back-to-back DADDs carry a stall of 1)
    seg_c = {}
    for r in [outer] + inner:
        t, free = 0, 0
        own = [k for k in range(r[0], r[1] + 1) if not any(q != r and q in inner and q[0] <= k <= q[1] for q in ([] if r != outer else inner))]
        for k in own:
            issue = max(t, free) if is_fp64(ins[k]) else t
            if is_fp64(ins[k]):
                free = issue + a.fp64_issue
            ins[k]["at"] = issue
            t = issue + cost(ins[k])
        seg_c[r] = t
    for k in range(outer[0], outer[1] + 1):
        depth = sum(1 for r in inner if r[0] <= k <= r[1])
        w = a.trip ** depth
        total_n += 0 if ins[k].get("skipped") else w
        e = by_line.setdefault(ins[k]["where"], [0.0, 0.0])
        e[0] += w * cost(ins[k]); e[1] += w
        if is_fp64(ins[k]):
            fp64 += w
        if not ins[k].get("skipped"):
            o = opcode(ins[k]["text"]).split(".")[0]
            mix[o] = mix.get(o, 0) + w
    total_c = seg_c[outer] + sum(a.trip * seg_c[r] for r in inner)
    s = summarize(ins, *outer)
    print("step loop [%#x..%#x]: %d static instructions, %d static stall cycles, %d with scoreboard waits" % (ins[outer[0]]["addr"], ins[outer[1]]["addr"], s["n"], s["cycles"], s["waits"]))
    for r in sorted(inner):
        t = summarize(ins, *r)
        print("  inner loop [%#x..%#x]: %d instructions, %d stall cycles per trip (%d with the FP64 pipe); DFMA %d DMUL %d DADD %d LDL %d STL %d" % (
            ins[r[0]]["addr"], ins[r[1]]["addr"], t["n"], t["cycles"], seg_c[r], t["ops"].get("DFMA", 0), t["ops"].get("DMUL", 0), t["ops"].get("DADD", 0),
            t["ops"].get("LDL", 0), t["ops"].get("STL", 0)))
    print("per step with %d trips per inner loop: %.0f instructions (%.0f FP64), %.0f cycles (stall fields + FP64 pipe occupancy); FP64 issue floor %.0f" % (a.trip, total_n, fp64, total_c, a.fp64_issue * fp64))
    fl = {k: mix.get(k, 0) for k in ("DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "MUFU")}
    print("arithmetic per step: " + ", ".join("%s %d" % kv for kv in fl.items() if kv[1]) + "; flops %d (FP64) %d (FP32)" % (
        2 * fl["DFMA"] + fl["DMUL"] + fl["DADD"], 2 * fl["FFMA"] + fl["FMUL"] + fl["FADD"]))
    for (f, ln), (c, n) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:a.lines]:
        print("  %6.0f cycles %5.0f instructions  %s:%d" % (c, n, f, ln))
    if a.phases:
        # outermost frame of the chain inside the given file = the call the step function makes; the frames below it
        # (callee side) refine it
        src = {}
        def text_of(f, ln):
            if f not in src:
                cands = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assistedmanipulation_b200")) if f in fs]
                src[f] = open(cands[0]).read().splitlines() if cands else []
            return src[f][ln - 1].strip()[:110] if 0 < ln <= len(src[f]) else ""
        tree = {}
        for k in range(outer[0], outer[1] + 1):
            it = ins[k]
            if it.get("skipped"):
                continue
            w = a.trip ** sum(1 for r in inner if r[0] <= k <= r[1])
            ch = it["chain"]
            idx = [j for j, fr in enumerate(ch) if fr[0] == a.phases]
            key = [("(outside " + a.phases + ")", 0)] if not idx else [ch[j] for j in range(idx[-1], max(idx[-1] - a.phase_depth, -1), -1)]
            node = tree
            for fr in key:
                node = node.setdefault(fr, dict(c=0.0, n=0.0, sub={}))
                node["c"] += w * cost(it); node["n"] += w
                node = node["sub"]
        def show(node, indent):
            for fr, v in sorted(node.items(), key=lambda kv: -kv[1]["n"]):
                if v["n"] < 20 and indent:
                    continue
                print("%s%6.0f instructions %6.0f stall cycles  %s:%d  %s" % ("  " * indent, v["n"], v["c"], fr[0], fr[1], text_of(*fr) if fr[1] else ""))
                show(v["sub"], indent + 1)
        print("step by call site (inline chains):")
        show(tree, 1)
    if a.dump:
        for k in range(outer[0], outer[1] + 1):
            it = ins[k]
            print("%06x @%-6d s%-2d %s w%02x b%d/%d  %-70s %s:%d" % (it["addr"], it.get("at", -1), it["stall"], "Y" if it["yield_"] else " ", it["wait"], it["wbar"], it["rbar"], it["text"], it["where"][0], it["where"][1]))


if __name__ == "__main__":
    sys.exit(main())
