"""B200-native MPPI rollout engine behind the reference's controller API (LuigiVan01/AssistedManipulation,
src/controller/mppi.hpp): abi = ctypes mirror of include/mppi_b200.h, engine.Engine = one Trajectory on the GPU,
forecast.DeviceForecast = the batched wrench forecast producer, formats = the reference's on-disk formats,
sharding = host-side protocol of the sharded rollout set."""
