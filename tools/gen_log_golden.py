"""Golden CSV logs for SURVEY §8f-4 written by the REFERENCE's own logger (logging/mppi.cpp + csv.hpp + file.hpp)
over the reference's own mppi::Trajectory — both compiled unmodified into oracle/_ref/libmppi_ref.so
(make -C oracle ref). Run here (needs /root/reference); the files travel under tests/golden/ref_logs/.

    python tools/gen_log_golden.py
"""
import ctypes as C
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import ref_lib  # noqa: E402

_dp = C.POINTER(C.c_double)


def main():
    ref = ref_lib.load()
    ref.ref_logger_create.argtypes = [C.c_char_p, C.c_uint, C.c_size_t]
    ref.ref_logger_create.restype = C.c_void_p
    ref.ref_logger_log.argtypes = [C.c_void_p, C.c_void_p]
    ref.ref_logger_destroy.argtypes = [C.c_void_p]
    out = os.path.join(ROOT, "tests", "golden", "ref_logs")
    shutil.rmtree(out, ignore_errors=True)
    for name, case in cases.LOG_CASES.items():
        holder = cases.config_for(case)
        params = case["params"]()
        h = ref.ref_create(C.byref(holder.cfg), C.cast(C.byref(params), C.c_void_p), C.sizeof(params))
        assert h
        folder = os.path.join(out, name)
        lg = ref.ref_logger_create(folder.encode(), holder.cfg.control_dof, case["K"] + 2)
        assert lg
        x0 = np.ascontiguousarray(case["x0"], dtype=np.float64)
        for u in range(case["updates"]):
            assert ref.ref_update(h, x0.ctypes.data_as(_dp), u * case["cadence"], None) == 0
            ref.ref_logger_log(lg, h)
            ref.ref_logger_log(lg, h)   # a second call at the same update time writes nothing (mppi.cpp:86-88)
        ref.ref_logger_destroy(lg)
        ref.ref_destroy(h)
        print(name, sorted(os.listdir(folder)))


if __name__ == "__main__":
    main()
