"""Python host side of the batched wrench forecast producer (C ABI mppi_b200_forecast_*): the device
counterpart of reference src/controller/forecast.hpp Forecast::update / forecast for `batch` forecasters at once."""
import ctypes as C

import numpy as np

from . import abi

_dp = C.POINTER(C.c_double)


class DeviceForecast:
    """The CUDA producer through the C ABI; `batch` forecasters, tables read back to the host."""

    def __init__(self, typ, hw, dt, order, batch=1, initial=None, device=0):
        self.lib = abi.load_library()
        cfg = abi.ForecastConfig(type=typ, batch=batch, device=device, order=order, time_step=dt,
                                 horison=hw, window=hw)
        self.batch = batch
        h = C.c_void_p()
        init = None if initial is None else np.ascontiguousarray(initial, dtype=np.float64)
        rc = self.lib.mppi_b200_forecast_create(C.byref(cfg), None if init is None else init.ctypes.data_as(_dp), C.byref(h))
        if rc != 0:
            raise RuntimeError("mppi_b200_forecast_create: %d %s" % (rc, self.lib.mppi_b200_forecast_last_error(None).decode()))
        self.h = h

    def update(self, m, t):
        m = np.ascontiguousarray(np.broadcast_to(np.asarray(m, dtype=np.float64).reshape(-1, 6), (self.batch, 6)))
        assert self.lib.mppi_b200_forecast_update(self.h, m.ctypes.data_as(_dp), t) == 0

    def update_time(self, t):
        assert self.lib.mppi_b200_forecast_update_time(self.h, t) == 0

    def table(self, t, dt, steps):
        out = np.zeros((self.batch, steps, 6))
        rc = self.lib.mppi_b200_forecast_table(self.h, t, dt, steps, out.ctypes.data_as(_dp))
        assert rc == 0, (rc, self.lib.mppi_b200_forecast_last_error(self.h).decode())
        return out

    def table_device(self, t, dt, steps):
        p = C.c_void_p()
        rc = self.lib.mppi_b200_forecast_table_device(self.h, t, dt, steps, C.byref(p))
        assert rc == 0, rc
        return p

    def forecast(self, t):
        return self.table(t, 1.0, 1)[0, 0]

    def close(self):
        if self.h:
            self.lib.mppi_b200_forecast_destroy(self.h)
            self.h = None
