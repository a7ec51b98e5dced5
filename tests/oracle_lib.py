"""ctypes loader for the CPU oracle (oracle/liboracle.so) — test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from assistedmanipulation_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_dp = C.POINTER(C.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def load():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_create.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_size_t]
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_update.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp]
    lib.oracle_update.restype = C.c_int
    lib.oracle_get.argtypes = [C.c_void_p, _dp, C.c_double]
    lib.oracle_set_optimal.argtypes = [C.c_void_p, _dp]
    lib.oracle_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.oracle_read.restype = C.c_int
    lib.oracle_query.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    lib.oracle_query.restype = C.c_int
    lib.oracle_phase_seconds.argtypes = [C.c_void_p, _dp]
    lib.oracle_sg_weights.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    for f in (lib.oracle_left_barrier, lib.oracle_right_barrier, lib.oracle_quadratic):
        f.argtypes = [C.c_double] * 4
        f.restype = C.c_double
    for f in (lib.oracle_upper_log_barrier, lib.oracle_lower_log_barrier):
        f.argtypes = [C.c_double] * 5
        f.restype = C.c_double
    lib.oracle_objective_probe.argtypes = [C.c_int, C.c_void_p, _dp, C.c_long, _dp]
    lib.oracle_tank.argtypes = [C.c_double, C.c_double, C.c_double, _dp]
    lib.oracle_sg_run.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int, _dp, C.c_double, _dp, _dp]
    lib.oracle_robot_fk.argtypes = [_dp, _dp, _dp, _dp]
    lib.oracle_robot_nle.argtypes = [_dp, _dp, _dp]
    lib.oracle_robot_aba.argtypes = [_dp, _dp, _dp, _dp]
    lib.oracle_robot_crba.argtypes = [_dp, _dp]
    lib.oracle_robot_kinematics.argtypes = [_dp] * 6
    lib.oracle_robot_rollout.argtypes = [_dp, _dp, C.c_int, C.c_double, _dp]
    lib.oracle_count_step_flops.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.oracle_count_step_flops.restype = C.c_uint64
    return lib


def ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class Oracle:
    """Thin handle around oracle_create/update/read, same call shapes as tests' Engine wrapper."""

    def __init__(self, lib, holder, objective):
        self.lib = lib
        self.holder = holder
        self.h = lib.oracle_create(C.byref(holder.cfg), C.cast(C.byref(objective), C.c_void_p), C.sizeof(objective))
        if not self.h:
            raise RuntimeError(lib.oracle_last_error().decode())

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def query(self, what):
        v = C.c_int64()
        assert self.lib.oracle_query(self.h, what, C.byref(v)) == 0
        return v.value

    def update(self, state, time, wrench=None, noise=None):
        state = np.ascontiguousarray(state, dtype=np.float64)
        wrench = None if wrench is None else np.ascontiguousarray(wrench, dtype=np.float64)
        noise = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        return self.lib.oracle_update(self.h, ptr(state), time, ptr(wrench), ptr(noise))

    def read(self, what, count, dtype=np.float64):
        out = np.zeros(count, dtype=dtype)
        rc = self.lib.oracle_read(self.h, what, out.ctypes.data_as(C.c_void_p), out.nbytes)
        assert rc == 0, rc
        return out

    def set_optimal(self, U):
        """test hook: continue from this published control sequence (see mppi_oracle.hpp set_optimal)"""
        U = np.ascontiguousarray(U, dtype=np.float64)
        self.lib.oracle_set_optimal(self.h, ptr(U))

    def get(self, time):
        nu = self.query(abi.QUERY_CONTROL_DOF)
        out = np.zeros(nu)
        self.lib.oracle_get(self.h, ptr(out), time)
        return out
