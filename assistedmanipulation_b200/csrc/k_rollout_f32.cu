// FP32 instantiations of the rollout kernel (fast mode; BASELINE.json configs 3-5).
#include "kernels.cuh"
#include "model_init.h"
namespace mppi_b200 {
__constant__ RobotModel<float> c_model_f32;
__constant__ FastModel<float> c_fast_f32;
__constant__ FastModel<double> c_fast_f32_state;   // the FP64 state path of the fast mode (rollout_core.cuh MIXED_SOLVER)
}
#define MPPI_DEVICE_MODEL c_model_f32
#define MPPI_DEVICE_FAST_MODEL c_fast_f32
#define MPPI_DEVICE_FAST_MODEL64 (&c_fast_f32_state)
#define MPPI_ROLLOUT_F32 1
#include "k_rollout.cuh"
namespace mppi_b200 {
cudaError_t upload_robot_model_f32() {
    const RobotModel<float> m = make_robot_model<float>();
    cudaError_t e = cudaMemcpyToSymbol(c_model_f32, &m, sizeof m);
    if (e != cudaSuccess) return e;
    const FastModel<float> f = make_fast_model<float>();
    e = cudaMemcpyToSymbol(c_fast_f32, &f, sizeof f);
    if (e != cudaSuccess) return e;
    const FastModel<double> fd = make_fast_model<double>();
    return cudaMemcpyToSymbol(c_fast_f32_state, &fd, sizeof fd);
}
cudaError_t launch_rollout_f32(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query) {
    return launch_rollout_r<float>(d, variant, faithful, params, optimal_only, s, chase_query);
}
}  // namespace mppi_b200
