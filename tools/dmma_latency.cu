// DMMA.8x8x4 issue/latency micro-benchmark for B200 (sm_100a): cycles per DMMA per scheduler as a function of the number
// of INDEPENDENT accumulator chains per warp and of the warps per scheduler.  One chain = the dependent-issue latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void __launch_bounds__(512) k(double* out, long long* cyc, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CH>
static void run(double* out, long long* cyc, int sms, int warps) {
    const int iters = 4000;
    k<CH><<<sms, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    k<CH><<<sms, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    const double per_warp = (double)h / (iters * 8.0 * CH);                 // cycles between DMMA issues of one warp
    const double per_sched = per_warp / (warps / 4.0);                      // cycles per DMMA on one scheduler's pipe
    printf("chains %d  warps/scheduler %d : %.1f cycles per DMMA per warp, %.1f per scheduler (16 = pipe saturated)\n", CH, warps / 4, per_warp, per_sched);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 512)); CK(cudaMalloc(&cyc, 8));
    for (int warps = 4; warps <= 16; warps += 4) {
        run<1>(out, cyc, p.multiProcessorCount, warps);
        run<2>(out, cyc, p.multiProcessorCount, warps);
        run<3>(out, cyc, p.multiProcessorCount, warps);
        run<4>(out, cyc, p.multiProcessorCount, warps);
        run<6>(out, cyc, p.multiProcessorCount, warps);
        run<8>(out, cyc, p.multiProcessorCount, warps);
    }
    return 0;
}
