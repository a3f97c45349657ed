// C-ABI glue: error text, host-buffer wrappers (H2D, launch, D2H, optional CUDA-event timing), FP64 peak probe.
#include <stdarg.h>
#include <mutex>
#include "common.cuh"

namespace wb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// One warp: lane l adds blocks l, l + 32, ... in index order (double-double), then a fixed xor-tree over the lanes.
// Deterministic for a given grid; replaces a single-thread loop that cost 25 us for 128 partials (r02 launch list).
__global__ void final_reduce_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out4) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += 32) {
        dd_add_dd(re, dd{partials[b * 4 + 0], partials[b * 4 + 1]});
        dd_add_dd(im, dd{partials[b * 4 + 2], partials[b * 4 + 3]});
    }
    re = warp_reduce_dd(re);
    im = warp_reduce_dd(im);
    if (threadIdx.x == 0) { out4[0] = re.hi; out4[1] = re.lo; out4[2] = im.hi; out4[3] = im.lo; }
}

int device_sm_count(int device, int* sms) {
    static int cache[64];
    static std::mutex mu;
    if (device < 0 || device >= 64) { set_error("bad device ordinal %d", device); return WB200_EINVAL; }
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[device]) {
        int v = 0;
        WB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        cache[device] = v;
    }
    *sms = cache[device];
    return WB200_OK;
}

// ---- caching device allocator ------------------------------------------------------------------
namespace {
struct PoolBlock {
    void* p;
    size_t bytes;
    int device;
    bool busy;
};
struct Pool {
    PoolBlock blk[64];
    int n = 0;
    ~Pool() {                                  // thread exit: give idle blocks back (busy ones belong to a live call)
        for (int i = 0; i < n; ++i)
            if (blk[i].p && !blk[i].busy) {
                int cur = 0;
                if (cudaGetDevice(&cur) != cudaSuccess) break;      // runtime already torn down
                if (blk[i].device != cur) cudaSetDevice(blk[i].device);
                cudaFree(blk[i].p);
                if (blk[i].device != cur) cudaSetDevice(cur);
                blk[i] = PoolBlock{nullptr, 0, -1, false};
            }
    }
};
thread_local Pool g_pool;
}  // namespace

int pool_alloc(void** p, size_t bytes) {
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (bytes < 256) bytes = 256;
    Pool& pl = g_pool;
    int best = -1;
    for (int i = 0; i < pl.n; ++i) {
        const PoolBlock& b = pl.blk[i];
        if (b.busy || !b.p || b.device != dev || b.bytes < bytes) continue;
        if (best < 0 || b.bytes < pl.blk[best].bytes) best = i;
    }
    if (best >= 0 && pl.blk[best].bytes <= 4 * bytes + (1u << 20)) {   // do not pin a huge block under a tiny request
        pl.blk[best].busy = true;
        *p = pl.blk[best].p;
        return WB200_OK;
    }
    // round up so that slowly growing requests (one more mode per sampler step) keep hitting the same block
    size_t want = 256;
    while (want < bytes) want <<= 1;
    if (want > (64u << 20)) want = (bytes + (16u << 20) - 1) / (16u << 20) * (16u << 20);
    int slot = -1;
    for (int i = 0; i < pl.n && slot < 0; ++i)                             // an invalidated entry (failed cudaMalloc earlier)
        if (!pl.blk[i].p) slot = i;
    if (slot >= 0) {
    } else if (pl.n < 64) {
        slot = pl.n++;
        pl.blk[slot] = PoolBlock{nullptr, 0, -1, false};
    } else {
        for (int i = 0; i < pl.n; ++i)                                     // table full: recycle the smallest idle block
            if (!pl.blk[i].busy && (slot < 0 || pl.blk[i].bytes < pl.blk[slot].bytes)) slot = i;
        if (slot >= 0) {
            int cur = dev;
            if (pl.blk[slot].device != cur) cudaSetDevice(pl.blk[slot].device);
            cudaFree(pl.blk[slot].p);
            if (pl.blk[slot].device != cur) cudaSetDevice(cur);
            pl.blk[slot] = PoolBlock{nullptr, 0, -1, false};             // never leave a freed pointer in the table
        }
    }
    void* q = nullptr;
    WB_CUDA(cudaMalloc(&q, want));
    if (slot < 0) {                                                        // every block is in use: untracked allocation
        *p = q;
        return WB200_OK;
    }
    pl.blk[slot] = PoolBlock{q, want, dev, true};
    *p = q;
    return WB200_OK;
}

int stream_alloc(void** p, size_t bytes, cudaStream_t st) {
    static bool configured[64];
    static std::mutex mu;
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lk(mu);
        if (!configured[dev]) {      // keep freed blocks in the pool across synchronisations (default threshold 0 returns them to the OS)
            cudaMemPool_t pool;
            WB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            unsigned long long keep = ~0ull;
            WB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            configured[dev] = true;
        }
    }
    if (bytes < 256) bytes = 256;
    WB_CUDA(cudaMallocAsync(p, bytes, st));
    return WB200_OK;
}

void pool_free(void* p) {
    Pool& pl = g_pool;
    for (int i = 0; i < pl.n; ++i)
        if (pl.blk[i].p == p) { pl.blk[i].busy = false; return; }
    cudaFree(p);   // untracked (table was full)
}

int check_perm_args(int n, int method, uint64_t k0, uint64_t k1);
int perm_f64_dev(const double*, int, int, uint64_t, uint64_t, double*, void*, cudaStream_t);
int perm_i64_dev(const int64_t*, int, int, uint64_t, uint64_t, unsigned long long*, cudaStream_t);

// RAII device scratch for the host-buffer wrappers
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) pool_free(p); }
    int alloc(size_t bytes) { return pool_alloc(&p, bytes); }
};
struct Timer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Timer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    // start(st, want = false) / stop(st, nullptr) are no-ops: the host wrappers pay for events only when the caller asks for
    // the kernel time (the end-to-end path does not)
    int start(cudaStream_t st, bool want = true) {
        if (!want) return 0;
        WB_CUDA(cudaEventCreate(&e0)); WB_CUDA(cudaEventCreate(&e1)); WB_CUDA(cudaEventRecord(e0, st)); return 0;
    }
    int stop(cudaStream_t st, double* ms) {
        if (!e0) return 0;
        WB_CUDA(cudaEventRecord(e1, st));
        WB_CUDA(cudaEventSynchronize(e1));
        float f = 0;
        WB_CUDA(cudaEventElapsedTime(&f, e0, e1));
        if (ms) *ms = f;
        return 0;
    }
};

// ---- FP64 peak probe ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_dfma(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_dmma(double* out, int iters, double a, double b) {
    double c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace wb

using namespace wb;

extern "C" const char* wb200_last_error(void) { return g_err; }
extern "C" int wb200_version(void) { return 100; }
extern "C" int wb200_device_count(int* count) {
    if (!count) return WB200_EINVAL;
    WB_CUDA(cudaGetDeviceCount(count));
    return WB200_OK;
}

extern "C" int wb200_fp64_peak(int device, int kind, double* tflops) {
    if (!tflops) return WB200_EINVAL;
    WB_CUDA(cudaSetDevice(device));
    int sms = 0;
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    const int grid = sms * 4, threads = 256, iters = 20000;
    DevBuf out;
    if (out.alloc(sizeof(double) * grid * threads)) return WB200_ECUDA;
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        Timer tm;
        if (tm.start(0)) return WB200_ECUDA;
        if (kind == 0) peak_dfma<<<grid, threads>>>((double*)out.p, iters, 1.0000001, 1e-9);
        else peak_dmma<<<grid, threads>>>((double*)out.p, iters, 1.0000001, 1e-9);
        double ms = 0;
        if (tm.stop(0, &ms)) return WB200_ECUDA;
        if (rep >= 1 && ms < best) best = ms;
    }
    WB_CUDA(cudaGetLastError());
    const double flops = kind == 0 ? 2.0 * 8 * 8 * iters * (double)grid * threads
                                   : 2.0 * 256 * 4 * 8 * iters * (double)grid * (threads / 32);
    *tflops = flops / (best * 1e-3) * 1e-12;
    return WB200_OK;
}

extern "C" int wb200_hafnian_host(int device, const double* A, const double* D, int n, uint64_t j0, uint64_t j1,
                                  double out4[4], double* kernel_ms) {
    if (!A || !out4) { set_error("hafnian: null pointer"); return WB200_EINVAL; }
    if (n < 2 || (n & 1)) { set_error("hafnian: n must be even and >= 2 (got %d)", n); return WB200_EINVAL; }
    if (n > 64) { set_error("hafnian: n = %d exceeds the DMMA kernel limit of 64", n); return WB200_ENOSUP; }
    WB_CUDA(cudaSetDevice(device));
    DevBuf dA, dD, dout, ws;
    const size_t wsb = wb200_hafnian_workspace_bytes(n);
    if (dA.alloc(sizeof(double) * 2 * n * n) || dout.alloc(4 * sizeof(double)) || ws.alloc(wsb)) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(dA.p, A, sizeof(double) * 2 * n * n, cudaMemcpyHostToDevice));
    if (D) {
        if (dD.alloc(sizeof(double) * 2 * n)) return WB200_ECUDA;
        WB_CUDA(cudaMemcpy(dD.p, D, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    }
    Timer tm;
    if (tm.start(0, kernel_ms != nullptr)) return WB200_ECUDA;
    int rc = wb200_hafnian_dev((const double*)dA.p, (const double*)dD.p, n, j0, j1, (double*)dout.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    if (tm.stop(0, kernel_ms)) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(out4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return WB200_OK;
}

extern "C" int wb200_perm_host(int device, const double* M, int n, int method, uint64_t k0, uint64_t k1,
                               double out4[4], double* kernel_ms) {
    if (!M || !out4) { set_error("perm: null pointer"); return WB200_EINVAL; }
    { int rc0 = check_perm_args(n, method, k0, k1); if (rc0) return rc0; }
    WB_CUDA(cudaSetDevice(device));
    DevBuf dM, dout, ws;
    const size_t wsb = wb200_perm_workspace_bytes(n);
    if (dM.alloc(sizeof(double) * 2 * n * n) || dout.alloc(4 * sizeof(double)) || ws.alloc(wsb)) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(dM.p, M, sizeof(double) * 2 * n * n, cudaMemcpyHostToDevice));
    Timer tm;
    if (tm.start(0, kernel_ms != nullptr)) return WB200_ECUDA;
    int rc = wb200_perm_dev((const double*)dM.p, n, method, k0, k1, (double*)dout.p, ws.p, wsb, nullptr);
    if (rc) return rc;
    if (tm.stop(0, kernel_ms)) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(out4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return WB200_OK;
}

extern "C" int wb200_perm_f64_host(int device, const double* M, int n, int method, uint64_t k0, uint64_t k1,
                                   double out2[2], double* kernel_ms) {
    if (!M || !out2) { set_error("perm: null pointer"); return WB200_EINVAL; }
    { int rc0 = check_perm_args(n, method, k0, k1); if (rc0) return rc0; }
    WB_CUDA(cudaSetDevice(device));
    DevBuf dM, dout, ws;
    if (dM.alloc(sizeof(double) * n * n) || dout.alloc(4 * sizeof(double)) || ws.alloc(wb200_perm_workspace_bytes(n))) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(dM.p, M, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    Timer tm;
    if (tm.start(0, kernel_ms != nullptr)) return WB200_ECUDA;
    int rc = perm_f64_dev((const double*)dM.p, n, method, k0, k1, (double*)dout.p, ws.p, nullptr);
    if (rc) return rc;
    if (tm.stop(0, kernel_ms)) return WB200_ECUDA;
    double o4[4];
    WB_CUDA(cudaMemcpy(o4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    out2[0] = o4[0]; out2[1] = o4[1];
    return WB200_OK;
}

extern "C" int wb200_perm_int64_host(int device, const int64_t* M, int n, int method, uint64_t k0, uint64_t k1,
                                     int64_t* out, double* kernel_ms) {
    if (!M || !out) { set_error("perm: null pointer"); return WB200_EINVAL; }
    { int rc0 = check_perm_args(n, method, k0, k1); if (rc0) return rc0; }
    WB_CUDA(cudaSetDevice(device));
    DevBuf dM, dout;
    if (dM.alloc(sizeof(int64_t) * n * n) || dout.alloc(sizeof(int64_t))) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(dM.p, M, sizeof(int64_t) * n * n, cudaMemcpyHostToDevice));
    Timer tm;
    if (tm.start(0, kernel_ms != nullptr)) return WB200_ECUDA;
    int rc = perm_i64_dev((const int64_t*)dM.p, n, method, k0, k1, (unsigned long long*)dout.p, nullptr);
    if (rc) return rc;
    if (tm.stop(0, kernel_ms)) return WB200_ECUDA;
    WB_CUDA(cudaMemcpy(out, dout.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
    return WB200_OK;
}

extern "C" int wb200_release_scratch(void) {
    Pool& pl = g_pool;
    int cur = 0;
    cudaGetDevice(&cur);
    int kept = 0;
    for (int i = 0; i < pl.n; ++i) {
        if (pl.blk[i].busy) { pl.blk[kept++] = pl.blk[i]; continue; }
        if (!pl.blk[i].p) continue;
        cudaSetDevice(pl.blk[i].device);
        cudaFree(pl.blk[i].p);
    }
    for (int i = kept; i < pl.n; ++i) pl.blk[i] = PoolBlock{nullptr, 0, -1, false};   // no freed pointers in the tail
    pl.n = kept;
    cudaSetDevice(cur);
    return WB200_OK;
}
