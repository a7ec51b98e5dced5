// FP32 instantiations of the rollout kernel (fast mode; BASELINE.json configs 3-5).
#include "kernels.cuh"
#include "model_init.h"
namespace mppi_b200 {
__constant__ RobotModel<float> c_model_f32;
}
#define MPPI_DEVICE_MODEL c_model_f32
#include "k_rollout.cuh"
namespace mppi_b200 {
cudaError_t upload_robot_model_f32() {
    const RobotModel<float> m = make_robot_model<float>();
    return cudaMemcpyToSymbol(c_model_f32, &m, sizeof m);
}
cudaError_t launch_rollout_f32(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s) {
    return launch_rollout_r<float>(d, variant, faithful, params, optimal_only, s);
}
}  // namespace mppi_b200
