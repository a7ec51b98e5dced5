"""tools/sass_cycles.py — the static issue-cycle model of DESIGN.md section 5 — run on the objects build() leaves in
csrc/build (nvcc cross-compiles here; no GPU involved). Pins the numbers the documentation quotes to a band, so a
change that silently loses the kernel's schedule (the inertia loop's is fragile) fails a CPU test."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "assistedmanipulation_b200", "csrc", "build")


def _model(obj, pattern, *extra):
    path = os.path.join(OBJ, obj)
    if not os.path.exists(path) or not shutil.which("nvdisasm") or not shutil.which("cuobjdump"):
        pytest.skip("needs the built objects and the CUDA binary utilities")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_cycles.py"), path, pattern, *extra], check=True, capture_output=True, text=True).stdout
    m = re.search(r"per step with 7 trips per inner loop: (\d+) instructions \((\d+) FP64\), (\d+) cycles .*floor (\d+)", out)
    assert m, out
    return tuple(int(v) for v in m.groups()), out


def test_config2_kernel_stays_near_its_fp64_issue_floor():
    (instr, fp64, cycles, floor), out = _model("k_rollout_f64.o", "Li1ELb0ENS_11TrackPointPIdEELb0ELb0EEE")
    assert floor == 2 * fp64
    assert 2000 <= fp64 <= 2200, out          # 2585 before the round's second half, 2173 before the joint offsets' structural zeros
    assert instr <= 3100 and cycles <= 5800, out   # 3668 instructions / 7597 cycles before; 5250 with the guarded slow path counted; this loop-body build only runs behind MPPI_B200_BIG_FROM (A/B), its schedule moves by ~100 cycles with unrelated edits
    assert cycles <= 1.35 * floor, out
    assert "1 loops" not in out and "inner loop" in out     # the inertia pass is a loop body in this build


def test_unrolled_and_assisted_kernels_keep_their_instruction_counts():
    (instr, fp64, _, _), out = _model("k_rollout_f64.o", "Li1ELb0ENS_11TrackPointPIdEELb1ELb0EEE")
    assert instr <= 2100 and fp64 <= 1650, out     # 3120 / 2515 at first, 2445 / 1956 before the joint offsets' structural zeros
    # config 3 / 5 kernel: FP32 kinematics / RNEA / objective around the FP64 state path (solver, sines, integration) since round 2
    (instr, fp64, cycles, _), out = _model("k_rollout_f32.o", "IfLi4ELb0ENS_9AssistedPIfEELb0ELb0EEE")
    assert instr <= 5400 and 1300 <= fp64 <= 1600 and cycles <= 9200, out   # all FP32: 8832 instructions at first, 4705 at the end of round 1; 5238 (1456 FP64) with the FP64 state path
