// FP64 pipe micro-benchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
// Decides whether the hafnian power-trace products go to DMMA or the DFMA pipe (BASELINE.json north_star).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], double b0, double b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// CH independent accumulator tiles per warp, 8 mma per chain per iteration
template <int CH>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma1684(c[i], a, b, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[CH][4];
    double av[4] = {a, b, a, b};
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma1688(c[i], av, b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[CH][4];
    double av[8] = {a, b, a, b, a, b, a, b};
    double bv[4] = {b, a, b, a};
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma16816(c[i], av, bv);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA chains and DFMA chains in the same warp (ratio 1 dmma884 : NF dfma)
template <int CH, int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double c0[CH], c1[CH], f[CH * NF + 1];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < CH * NF; ++i) f[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                dmma884(c0[i], c1[i], a, b);
#pragma unroll
                for (int j = 0; j < NF; ++j) f[i * NF + j] = fma(f[i * NF + j], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < CH * NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA with one LDS.64 B-fragment fetch per mma (realistic operand streaming from smem)
template <int CH>
__global__ void __launch_bounds__(256) k_dmma884_lds(double* out, int iters, double a) {
    __shared__ double tab[64 * 32];
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) tab[i] = 1e-3 * (i % 7);
    __syncthreads();
    double c0[CH], c1[CH];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double b = tab[((it * 8 + r) & 63) * 32 + lane];
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a + i, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best * 1e-3;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    const int iters = 20000;
    for (int bps = 1; bps <= 8; bps *= 2) {
        for (int threads = 128; threads <= 256; threads *= 2) {
        int grid = sms * bps; int warps = threads / 32;
        double t, fl;
        t = timeit([&] { k_dfma<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 8 * 8 * (double)iters * grid * threads;
        printf("dfma ch8        bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma884<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 256 * 4 * 8 * (double)iters * grid * warps;
        printf("dmma884 ch4     bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma884<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 256 * 8 * 8 * (double)iters * grid * warps;
        printf("dmma884 ch8     bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma1684<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 512 * 4 * 8 * (double)iters * grid * warps;
        printf("dmma1684 ch4    bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma1688<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 1024 * 4 * 8 * (double)iters * grid * warps;
        printf("dmma1688 ch4    bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma16816<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 2048 * 4 * 8 * (double)iters * grid * warps;
        printf("dmma16816 ch4   bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_mixed<4, 4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * (256 + 4 * 32) * 4 * 8 * (double)iters * grid * warps;
        printf("mixed 1mma:4fma bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_mixed<4, 8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * (256 + 8 * 32) * 4 * 8 * (double)iters * grid * warps;
        printf("mixed 1mma:8fma bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma884_lds<4><<<grid, threads>>>(out, iters, 1.0000001); });
        fl = 2.0 * 256 * 4 * 8 * (double)iters * grid * warps;
        printf("dmma884+lds/4   bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        t = timeit([&] { k_dmma884_lds<1><<<grid, threads>>>(out, iters, 1.0000001); });
        fl = 2.0 * 256 * 1 * 8 * (double)iters * grid * warps;
        printf("dmma884+lds/1   bps %d thr %d : %.3f ms  %.2f TFLOP/s\n", bps, threads, t * 1e3, fl / t * 1e-12);
        }
    }
    return 0;
}
