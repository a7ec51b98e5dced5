// Microbenchmark: how fast can ONE warp issue FP64 FMAs on a B200 SM sub-partition, as a function of the
// number of independent chains (ILP) and of the active-lane count? Answers whether the latency-bound rollout
// kernel (one warp per SM at K=4096) can gain from narrower warps or only from more warps per rollout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_issue dfma_issue.cu && ./dfma_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, long long *cycles, int active, int iters, double a, double b) {
    if ((int)(threadIdx.x & 31) >= active) return;
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP> void run(int warps, int active, double *out, long long *cyc) {
    const int iters = 2000;
    k<ILP><<<1, 32 * warps>>>(out, cyc, active, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    k<ILP><<<1, 32 * warps>>>(out, cyc, active, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %d active %2d ILP %2d : %.2f cycles per DFMA instruction per warp\n", warps, active, ILP, (double)c / (iters * 8.0 * ILP));
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int warps : {1, 4, 8}) for (int active : {32, 16, 8}) {
        run<1>(warps, active, out, cyc); run<2>(warps, active, out, cyc); run<4>(warps, active, out, cyc); run<8>(warps, active, out, cyc); run<16>(warps, active, out, cyc);
    }
    return 0;
}
