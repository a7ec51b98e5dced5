// FP64 instantiations of the rollout kernel (exact mode; BASELINE.json config 2).
#include "kernels.cuh"
#include "model_init.h"
namespace mppi_b200 {
__constant__ RobotModel<double> c_model_f64;
__constant__ FastModel<double> c_fast_f64;
}
#define MPPI_DEVICE_MODEL c_model_f64
#define MPPI_DEVICE_FAST_MODEL c_fast_f64
#define MPPI_DEVICE_FAST_MODEL64 nullptr
#include "k_rollout.cuh"
namespace mppi_b200 {
cudaError_t upload_robot_model_f64() {
    const RobotModel<double> m = make_robot_model<double>();
    cudaError_t e = cudaMemcpyToSymbol(c_model_f64, &m, sizeof m);
    if (e != cudaSuccess) return e;
    const FastModel<double> f = make_fast_model<double>();
    return cudaMemcpyToSymbol(c_fast_f64, &f, sizeof f);
}
cudaError_t launch_rollout_f64(const DeviceState &d, int variant, bool faithful, const void *params, bool optimal_only, cudaStream_t s, int *chase_query) {
    return launch_rollout_r<double>(d, variant, faithful, params, optimal_only, s, chase_query);
}
}  // namespace mppi_b200
