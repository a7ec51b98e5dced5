"""ctypes loader for oracle/_ref/libmppi_ref.so — the reference's own mppi.cpp / filter.cpp /
gaussian.hpp / gram_savitzky_golay.cpp, its objectives (track_point.cpp, assisted_manipulation.cpp, cost.hpp,
energy.hpp), frankaridgeback/dynamics.cpp and state.cpp compiled unmodified (oracle/Makefile `ref`). Only present
where /root/reference is mounted or the built library travelled with the snapshot."""
import ctypes as C
import os

from assistedmanipulation_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libmppi_ref.so")
_dp = C.POINTER(C.c_double)


def available():
    return os.path.exists(PATH)


def load():
    ref = C.CDLL(PATH)
    ref.ref_create.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.c_size_t]
    ref.ref_create.restype = C.c_void_p
    ref.ref_destroy.argtypes = [C.c_void_p]
    ref.ref_update.argtypes = [C.c_void_p, _dp, C.c_double, _dp]
    ref.ref_read.argtypes = [C.c_void_p, C.c_int, _dp, C.c_size_t]
    ref.ref_get.argtypes = [C.c_void_p, _dp, C.c_double]
    ref.ref_sg_weights.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    ref.ref_sg_run.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int, _dp, C.c_double, _dp, _dp]
    ref.ref_objective_probe.argtypes = [C.c_int, C.c_void_p, _dp, C.c_long, _dp]
    ref.ref_cost_functor.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, C.c_long, _dp]
    ref.ref_energy_tank.argtypes = [C.c_double, _dp, C.c_double, C.c_long, _dp]
    ref.ref_make_state.argtypes = [C.c_int, _dp]
    return ref
