// Permanent by Gray-code subset sums: BBFG/Glynn (thewalrus/_permanent.py:130-168) and Ryser (:86-127).
//
// Step k of the reference loops evaluates the Gray code g(k) = k ^ (k >> 1) with sign (-1)^k on the
// column sums r_c = sum_rows delta_row M[row, c]:
//   bbfg : delta = +1 / -1 for bit clear / set, k in [0, 2^(n-1)), result total / 2^(n-1)
//   ryser: r_c = -sum_{rows in g(k)} M[row, c],  k in [0, 2^n), result total (the sign absorbs (-1)^n)
// B200 mapping: every thread owns whole power-of-two aligned segments of the step index, keeps all n
// column sums in registers, seeds them from the Gray code of its first step (O(n^2)) and then applies the
// +-2*row (bbfg) / +-1*row (ryser) update per step.  Because segments are aligned and equally long, the
// flipped row ctz(k+1) is warp-uniform, so the row is one broadcast read from shared memory.  Products use
// four independent chains for FP64 ILP; per-thread Kahan sums feed a double-double block/grid reduction.
#include <type_traits>
#include "common.cuh"

namespace wb {

constexpr int PERM_THREADS = 256;

struct C128 {
    double re, im;
    __device__ __forceinline__ static C128 zero() { return {0.0, 0.0}; }
    __device__ __forceinline__ static C128 one() { return {1.0, 0.0}; }
};
__device__ __forceinline__ C128 cmul(C128 a, C128 b) {
    return {fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)};
}
__device__ __forceinline__ double cmul(double a, double b) { return a * b; }
__device__ __forceinline__ unsigned long long cmul(unsigned long long a, unsigned long long b) { return a * b; }

// x += s * y with s in {+-1, +-2}
__device__ __forceinline__ void axpy(C128& x, double s, C128 y) {
    x.re = fma(s, y.re, x.re);
    x.im = fma(s, y.im, x.im);
}
__device__ __forceinline__ void axpy(double& x, double s, double y) { x = fma(s, y, x); }
__device__ __forceinline__ void axpy(unsigned long long& x, long long s, unsigned long long y) {
    x += (unsigned long long)s * y;
}

template <typename S> struct Traits;
template <> struct Traits<C128> {
    using Coef = double;
    __device__ static C128 zero() { return {0.0, 0.0}; }
    __device__ static C128 one() { return {1.0, 0.0}; }
};
template <> struct Traits<double> {
    using Coef = double;
    __device__ static double zero() { return 0.0; }
    __device__ static double one() { return 1.0; }
};
template <> struct Traits<unsigned long long> {
    using Coef = long long;
    __device__ static unsigned long long zero() { return 0ull; }
    __device__ static unsigned long long one() { return 1ull; }
};

// Kahan accumulators
struct KahanC {
    double sr = 0, cr = 0, si = 0, ci = 0;
    __device__ __forceinline__ void add(C128 x, bool neg) {
        double xr = neg ? -x.re : x.re, xi = neg ? -x.im : x.im;
        double y = xr - cr, t = sr + y;
        cr = (t - sr) - y; sr = t;
        y = xi - ci; t = si + y;
        ci = (t - si) - y; si = t;
    }
    __device__ __forceinline__ cdd get() const { cdd o; o.re = {sr, -cr}; o.im = {si, -ci}; return o; }
};
struct KahanR {
    double s = 0, c = 0;
    __device__ __forceinline__ void add(double x, bool neg) {
        double y = (neg ? -x : x) - c, t = s + y;
        c = (t - s) - y; s = t;
    }
    __device__ __forceinline__ cdd get() const { cdd o; o.re = {s, -c}; o.im = {0.0, 0.0}; return o; }
};
struct AccI {
    unsigned long long s = 0;
    __device__ __forceinline__ void add(unsigned long long x, bool neg) { s += neg ? (0ull - x) : x; }
};
template <typename S> struct AccOf;
template <> struct AccOf<C128> { using type = KahanC; };
template <> struct AccOf<double> { using type = KahanR; };
template <> struct AccOf<unsigned long long> { using type = AccI; };

template <int NP, typename S>
__device__ __forceinline__ S product(const S (&r)[NP]) {
    S p0 = r[0], p1 = r[1], p2 = r[2], p3 = r[3];
#pragma unroll
    for (int c = 4; c < NP; c += 4) {
        p0 = cmul(p0, r[c]);
        p1 = cmul(p1, r[c + 1]);
        p2 = cmul(p2, r[c + 2]);
        p3 = cmul(p3, r[c + 3]);
    }
    return cmul(cmul(p0, p1), cmul(p2, p3));
}

// M in shared memory: smem[row * NP + col]; padded columns hold 0 and their sums are pinned to 1.
template <int NP, typename S>
__global__ void __launch_bounds__(PERM_THREADS)
perm_kernel(const S* __restrict__ Mg, int n, int ryser, uint64_t k0, uint64_t k1, int logL,
            double* __restrict__ partials, unsigned long long* __restrict__ iout) {
    using Coef = typename Traits<S>::Coef;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S* sM = reinterpret_cast<S*>(smem_raw);
    for (int i = threadIdx.x; i < n * NP; i += PERM_THREADS) {
        const int row = i / NP, col = i % NP;
        sM[i] = (col < n) ? Mg[row * n + col] : Traits<S>::zero();
    }
    __syncthreads();

    typename AccOf<S>::type acc;
    const uint64_t L = 1ull << logL;
    const uint64_t c_first = k0 >> logL, c_last = (k1 + L - 1) >> logL;  // chunk ids [c_first, c_last)
    const uint64_t nthreads = (uint64_t)gridDim.x * PERM_THREADS;
    const Coef step = ryser ? (Coef)1 : (Coef)2;
    for (uint64_t c = c_first + (uint64_t)blockIdx.x * PERM_THREADS + threadIdx.x; c < c_last; c += nthreads) {
        uint64_t kb = c << logL, ke = kb + L;
        if (kb < k0) kb = k0;
        if (ke > k1) ke = k1;
        if (kb >= ke) continue;
        // ---- seed column sums from g(kb)
        uint64_t gray = kb ^ (kb >> 1);
        S r[NP];
#pragma unroll
        for (int col = 0; col < NP; ++col) r[col] = Traits<S>::zero();
        for (int row = 0; row < n; ++row) {
            const bool set = (gray >> row) & 1ull;
            const Coef d = ryser ? (set ? (Coef)-1 : (Coef)0) : (set ? (Coef)-1 : (Coef)1);
            if (d != (Coef)0) {
#pragma unroll
                for (int col = 0; col < NP; ++col) axpy(r[col], d, sM[row * NP + col]);
            }
        }
#pragma unroll
        for (int col = 0; col < NP; ++col)
            if (col >= n) r[col] = Traits<S>::one();
        // ---- sweep
        for (uint64_t k = kb; k < ke; ++k) {
            acc.add(product<NP, S>(r), (k & 1ull) != 0);
            const uint64_t k1n = k + 1;
            const int row = __ffsll((long long)k1n) - 1;         // bit flipped between g(k) and g(k+1)
            if (row < n) {
                const bool set = ((k1n ^ (k1n >> 1)) >> row) & 1ull;  // new value of that bit
                const Coef d = set ? -step : step;
                const S* mrow = sM + row * NP;
#pragma unroll
                for (int col = 0; col < NP; ++col) axpy(r[col], d, mrow[col]);  // padded columns add 0
            }
        }
    }
    if constexpr (std::is_same<S, unsigned long long>::value) {
        unsigned long long v = acc.s;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(iout, v);
    } else {
        __shared__ double red[(PERM_THREADS / 32) * 4];
        block_reduce_store(acc.get(), red, partials);
    }
}

static int pick_logL(uint64_t total, uint64_t nthreads) {
    int logL = 6;
    while (logL < 20 && (total >> (logL + 1)) >= nthreads * 64) ++logL;
    return logL;
}

template <int NP, typename S>
static int launch_perm(const S* dM, int n, int method, uint64_t k0, uint64_t k1, double* partials,
                       unsigned long long* iout, int* grid_out, cudaStream_t st) {
    int dev = 0, sms = 0, occ = 1;
    WB_CUDA(cudaGetDevice(&dev));
    if (device_sm_count(dev, &sms)) return WB200_ECUDA;
    const size_t shm = sizeof(S) * (size_t)n * NP;
    WB_CUDA(cudaFuncSetAttribute(perm_kernel<NP, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, perm_kernel<NP, S>, PERM_THREADS, shm));
    if (occ < 1) occ = 1;
    const uint64_t total = k1 - k0;
    uint64_t maxgrid = (uint64_t)sms * occ;
    uint64_t want = (total + (uint64_t)PERM_THREADS * 64 - 1) / ((uint64_t)PERM_THREADS * 64);
    int grid = (int)(want < maxgrid ? (want ? want : 1) : maxgrid);
    const int logL = pick_logL(total, (uint64_t)grid * PERM_THREADS);
    perm_kernel<NP, S><<<grid, PERM_THREADS, shm, st>>>(dM, n, method, k0, k1, logL, partials, iout);
    WB_CUDA(cudaGetLastError());
    *grid_out = grid;
    return WB200_OK;
}

template <typename S>
static int dispatch_perm(const S* dM, int n, int method, uint64_t k0, uint64_t k1, double* partials,
                         unsigned long long* iout, int* grid, cudaStream_t st) {
    const int np = (n + 3) / 4;
    switch (np) {
        case 1: return launch_perm<4, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 2: return launch_perm<8, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 3: return launch_perm<12, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 4: return launch_perm<16, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 5: return launch_perm<20, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 6: return launch_perm<24, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 7: return launch_perm<28, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 8: return launch_perm<32, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 9: return launch_perm<36, S>(dM, n, method, k0, k1, partials, iout, grid, st);
        case 10: return launch_perm<40, S>(dM, n, method, k0, k1, partials, iout, grid, st);
    }
    set_error("perm: n = %d exceeds the register-resident kernel limit of 40", n);
    return WB200_ENOSUP;
}

constexpr int PERM_MAX_GRID = 4096;

static int check_perm_args(int n, int method, uint64_t k0, uint64_t k1) {
    if (n < 1 || n > 40) { set_error("perm: n = %d outside [1, 40]", n); return n > 40 ? WB200_ENOSUP : WB200_EINVAL; }
    if (method != 0 && method != 1) { set_error("perm: method must be 0 (bbfg) or 1 (ryser)"); return WB200_EINVAL; }
    const uint64_t steps = 1ull << (method ? n : n - 1);
    if (k0 > k1 || k1 > steps) { set_error("perm: bad step range"); return WB200_EINVAL; }
    return WB200_OK;
}

int perm_f64_dev(const double* dM, int n, int method, uint64_t k0, uint64_t k1, double* d_out4, void* ws,
                 cudaStream_t st) {
    int rc = check_perm_args(n, method, k0, k1);
    if (rc) return rc;
    int grid = 0;
    double* partials = reinterpret_cast<double*>(ws);
    rc = dispatch_perm<double>(dM, n, method, k0, k1, partials, nullptr, &grid, st);
    if (rc) return rc;
    final_reduce_kernel<<<1, 32, 0, st>>>(partials, grid, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

int perm_i64_dev(const int64_t* dM, int n, int method, uint64_t k0, uint64_t k1, unsigned long long* d_out,
                 cudaStream_t st) {
    int rc = check_perm_args(n, method, k0, k1);
    if (rc) return rc;
    int grid = 0;
    WB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), st));
    return dispatch_perm<unsigned long long>(reinterpret_cast<const unsigned long long*>(dM), n, method, k0, k1,
                                             nullptr, d_out, &grid, st);
}

}  // namespace wb

using namespace wb;

extern "C" size_t wb200_perm_workspace_bytes(int n) {
    (void)n;
    return sizeof(double) * (size_t)PERM_MAX_GRID * 4;
}

extern "C" int wb200_perm_dev(const double* dM, int n, int method, uint64_t k0, uint64_t k1, double* d_out4,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!dM || !d_out4 || !d_workspace) { set_error("perm: null pointer"); return WB200_EINVAL; }
    int rc = check_perm_args(n, method, k0, k1);
    if (rc) return rc;
    if (workspace_bytes < wb200_perm_workspace_bytes(n)) { set_error("perm: workspace too small"); return WB200_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int grid = 0;
    double* partials = reinterpret_cast<double*>(d_workspace);
    rc = dispatch_perm<C128>(reinterpret_cast<const C128*>(dM), n, method, k0, k1, partials, nullptr, &grid, st);
    if (rc) return rc;
    final_reduce_kernel<<<1, 32, 0, st>>>(partials, grid, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}
