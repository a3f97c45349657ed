// Shared device/host helpers for the walrus_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/walrus_b200.h"

namespace wb {

void set_error(const char* fmt, ...);

#define WB_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            wb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return WB200_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

// Caching device allocator behind the *_host wrappers (capi.cu): a thread-local free list per device, so a
// wrapper that is called in a loop (one call per mode of a sampler chain, one per pattern batch) pays for
// cudaMalloc / cudaFree only the first time.  Blocks return to the list when the wrapper's RAII buffers go out
// of scope, i.e. after its final synchronous copy; wb200_release_scratch() gives them back to the driver.
int pool_alloc(void** p, size_t bytes);
void pool_free(void* p);
#define WB_POOL(call)                       \
    do {                                    \
        int rc__ = (call);                  \
        if (rc__ != WB200_OK) return rc__;  \
    } while (0)

// Stream-ordered scratch for the *_dev entry points (cudaMallocAsync / cudaFreeAsync on the caller's stream): the call
// returns without synchronising and the block goes back to the device's memory pool when the stream reaches the free.
int stream_alloc(void** p, size_t bytes, cudaStream_t st);
struct StreamBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~StreamBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t bytes, cudaStream_t s) { st = s; return stream_alloc(&p, bytes, s); }
};

// ---------------------------------------------------------------------------------------------
// error-free transformations; compiled without fast-math so the compiler keeps them
// ---------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};

__host__ __device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

// acc += x in double-double (Neumaier-style renormalised accumulator)
__host__ __device__ __forceinline__ void dd_add(dd& acc, double x) {
    double s, e;
    two_sum(acc.hi, x, s, e);
    e += acc.lo;
    double hi = s + e;
    acc.lo = e - (hi - s);
    acc.hi = hi;
}

__host__ __device__ __forceinline__ void dd_add_dd(dd& acc, const dd& x) {
    double s, e;
    two_sum(acc.hi, x.hi, s, e);
    e += acc.lo + x.lo;
    double hi = s + e;
    acc.lo = e - (hi - s);
    acc.hi = hi;
}

struct cdd {  // complex double-double accumulator
    dd re, im;
};

__device__ __forceinline__ double shfl_xor_d(double v, int mask) {
    return __shfl_xor_sync(0xffffffffu, v, mask);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}

__device__ __forceinline__ dd warp_reduce_dd(dd v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        dd o;
        o.hi = shfl_xor_d(v.hi, off);
        o.lo = shfl_xor_d(v.lo, off);
        dd_add_dd(v, o);
    }
    return v;
}

// Block-level reduction of one complex double-double per thread into partials[blockIdx.x*4 .. +3].
// smem must hold 4 * (blockDim.x/32) doubles.
__device__ __forceinline__ void block_reduce_store(cdd acc, double* smem, double* partials) {
    acc.re = warp_reduce_dd(acc.re);
    acc.im = warp_reduce_dd(acc.im);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) {
        smem[warp * 4 + 0] = acc.re.hi;
        smem[warp * 4 + 1] = acc.re.lo;
        smem[warp * 4 + 2] = acc.im.hi;
        smem[warp * 4 + 3] = acc.im.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cdd t;
        t.re = {0.0, 0.0};
        t.im = {0.0, 0.0};
        for (int w = 0; w < nwarps; ++w) {
            dd_add_dd(t.re, dd{smem[w * 4 + 0], smem[w * 4 + 1]});
            dd_add_dd(t.im, dd{smem[w * 4 + 2], smem[w * 4 + 3]});
        }
        partials[blockIdx.x * 4 + 0] = t.re.hi;
        partials[blockIdx.x * 4 + 1] = t.re.lo;
        partials[blockIdx.x * 4 + 2] = t.im.hi;
        partials[blockIdx.x * 4 + 3] = t.im.lo;
    }
}

// Fixed-order final reduction of per-block partials -> out4 (deterministic for a given grid).
__global__ void final_reduce_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out4);

int device_sm_count(int device, int* sms);

}  // namespace wb
