// FP64 instantiations of the rollout kernel (exact mode; BASELINE.json config 2).
#include "kernels.cuh"
#include "model_init.h"
namespace mppi_b200 {
__constant__ RobotModel<double> c_model_f64;
}
#define MPPI_DEVICE_MODEL c_model_f64
#include "k_rollout.cuh"
namespace mppi_b200 {
cudaError_t upload_robot_model_f64() {
    const RobotModel<double> m = make_robot_model<double>();
    return cudaMemcpyToSymbol(c_model_f64, &m, sizeof m);
}
cudaError_t launch_rollout_f64(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s) {
    return launch_rollout_r<double>(d, variant, faithful, params, optimal_only, s);
}
}  // namespace mppi_b200
