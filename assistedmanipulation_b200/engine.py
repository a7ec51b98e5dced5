"""Python host side of the C ABI (include/mppi_b200.h): `Engine` = one mppi::Trajectory (or a batch of them) on the
GPU — create / update / get / read, the calls of reference src/controller/mppi.hpp:321-474. No CPU path: the shared
library must have been built (abi.load_library fails loudly otherwise)."""
import ctypes as C

import numpy as np

from . import abi

_dp = C.POINTER(C.c_double)
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = abi.load_library()
    return _LIB


def ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class Engine:
    def __init__(self, holder, objective):
        self.lib = lib()
        self.holder = holder
        h = C.c_void_p()
        rc = self.lib.mppi_b200_create(C.byref(holder.cfg), C.cast(C.byref(objective), C.c_void_p), C.sizeof(objective), C.byref(h))
        if rc != 0:
            raise RuntimeError("mppi_b200_create: %d %s" % (rc, self.lib.mppi_b200_last_error(None).decode()))
        self.h = h
        self._pointers, self._converted = {}, {}   # argument slot -> (array, its ctypes pointer): see _host_pointer

    def _host_pointer(self, slot, a):
        """Pointer to a host array of doubles. Building the ctypes pointer costs more than the rest of the call (2-3 us per
        argument against a 200 us update), so it is kept for as long as the caller passes the SAME array object — the usual
        control loop, which refills one state buffer in place. Arrays that need a conversion are converted on every call."""
        if a is None:
            return None
        held = self._pointers.get(slot)
        if held is not None and held[0] is a:
            return held[1]
        b = np.ascontiguousarray(a, dtype=np.float64)
        p = b.ctypes.data_as(_dp)
        if b is a:
            self._pointers[slot] = (a, p)
        else:
            self._pointers.pop(slot, None)
            self._converted[slot] = b   # keeps the converted copy alive for the duration of the call
        return p

    def close(self):
        if self.h:
            self.lib.mppi_b200_destroy(self.h)
            self.h = None

    def error(self):
        return self.lib.mppi_b200_last_error(self.h).decode()

    def query(self, what):
        v = C.c_int64()
        assert self.lib.mppi_b200_query(self.h, what, C.byref(v)) == 0
        return v.value

    def update(self, state, time, wrench=None, noise=None, seed=0, source=None):
        state_p, wrench_p = self._host_pointer("state", state), self._host_pointer("wrench", wrench)
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            source = abi.NOISE_HOST if source is None else source
            nptr = noise.ctypes.data_as(C.c_void_p)
        else:
            source, nptr = abi.NOISE_PHILOX, None
        return self.lib.mppi_b200_update(self.h, state_p, time, wrench_p, nptr, source, seed)

    def read(self, what, count, dtype=np.float64):
        out = np.zeros(count, dtype=dtype)
        rc = self.lib.mppi_b200_read(self.h, what, out.ctypes.data_as(C.c_void_p), out.nbytes)
        assert rc == 0, (rc, self.error())
        return out

    def get(self, time):
        out = np.zeros(self.query(abi.QUERY_CONTROL_DOF) * self.query(abi.QUERY_BATCH))
        assert self.lib.mppi_b200_get(self.h, ptr(out), time) == 0
        return out

    def device_seconds(self):
        s = C.c_double()
        self.lib.mppi_b200_last_update_device_seconds(self.h, C.byref(s))
        return s.value
