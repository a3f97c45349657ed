// Permanent by Gray-code subset sums: BBFG/Glynn (thewalrus/_permanent.py:130-168) and Ryser (:86-127).
//
// Step k of the reference loops evaluates the Gray code g(k) = k ^ (k >> 1) with sign (-1)^k on the
// column sums r_c = sum_rows delta_row M[row, c]:
//   bbfg : delta = +1 / -1 for bit clear / set, k in [0, 2^(n-1)), result total / 2^(n-1)
//   ryser: r_c = -sum_{rows in g(k)} M[row, c],  k in [0, 2^n), result total (the sign absorbs (-1)^n)
//
// B200 mapping.  The step index is cut in power-of-two aligned, equally long segments ("streams").  A stream
// is owned by SL adjacent lanes of a warp; lane `sub` keeps the column sums of the interleaved columns
// c * SL + sub (CPL of them) in registers.  SL = 1 (one thread holds all n sums) is the fastest shape up to
// n = 40 and is what runs there; SL = 4 with two streams per lane carries n = 41..64, where the sums no longer
// fit one thread (measured shapes: profiles/r01_perm_shapes.txt).
// Streams are seeded from the Gray code of their first step (O(n CPL) per lane) and then advance by the
// +-2*row (bbfg) / +-row (ryser) update; all streams of a warp are at the same offset of their segment, so
// the flipped row ctz(k) is warp-uniform and a lane's piece of the row is one LDS.128 shared by 32/SL lanes.
// Steps are taken in groups of G = max(SL, 2): every lane forms its partial product of each step of the group,
// then a transpose-reduce over the SL lanes (shuffle + one complex multiply per halving) leaves lane `sub`
// with the complete product of step k + sub, whose sign (-1)^sub is static.  There is no branch inside a
// group: steps outside [k0, k1) (ragged shard edges) are masked by a zero weight.  Terms are added in plain
// FP64 over 16 groups and then folded into a per-thread Kahan sum (the per-term rounding of the n-factor
// product dominates that block sum by an order of magnitude); per-thread sums feed a double-double
// warp -> block -> grid reduction in fixed order.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

namespace wb {

// One 256-thread CTA per SM.  With several smaller CTAs per SM (128 x 2, 64 x 4, 32 x 8 were measured) only 6.3 of
// the 8 warps the register file allows are resident on average and the FP64 pipe drops from 75 % to 71 % busy
// (profiles/r01_perm_block_shape.txt); -D overrides exist for that experiment (tools/gpu_perm_shape.sh).
#ifndef WB_PERM_THREADS
#define WB_PERM_THREADS 256
#define WB_PERM_MIN_CTAS 1
#endif
constexpr int PERM_THREADS = WB_PERM_THREADS;
constexpr int PERM_MIN_CTAS = WB_PERM_MIN_CTAS;   // <= 255 registers; the tile shapes below land at 8-16 warps per SM
constexpr int PERM_BLOCK_GROUPS = 16;   // groups added in plain FP64 before the compensated fold

struct C128 {
    double re, im;
};
__device__ __forceinline__ C128 cmul(C128 a, C128 b) {
    return {fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)};
}
__device__ __forceinline__ double cmul(double a, double b) { return a * b; }
__device__ __forceinline__ unsigned long long cmul(unsigned long long a, unsigned long long b) { return a * b; }
__device__ __forceinline__ C128 cadd(C128 a, C128 b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ double cadd(double a, double b) { return a + b; }
__device__ __forceinline__ unsigned long long cadd(unsigned long long a, unsigned long long b) { return a + b; }

// x += s * y with s in {0, +-1, +-2}
__device__ __forceinline__ void axpy(C128& x, double s, C128 y) {
    x.re = fma(s, y.re, x.re);
    x.im = fma(s, y.im, x.im);
}
__device__ __forceinline__ void axpy(double& x, double s, double y) { x = fma(s, y, x); }
__device__ __forceinline__ void axpy(unsigned long long& x, long long s, unsigned long long y) {
    x += (unsigned long long)s * y;
}

// shuffles among the lanes of one stream only (streams of a warp may run different numbers of groups)
__device__ __forceinline__ double shfl_xor_s(unsigned mask, double v, int m) { return __shfl_xor_sync(mask, v, m); }
__device__ __forceinline__ C128 shfl_xor_s(unsigned mask, C128 v, int m) {
    return {__shfl_xor_sync(mask, v.re, m), __shfl_xor_sync(mask, v.im, m)};
}
__device__ __forceinline__ unsigned long long shfl_xor_s(unsigned mask, unsigned long long v, int m) {
    return __shfl_xor_sync(mask, v, m);
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

// Per-thread accumulators, fed with the plain-FP64 sums of 16 groups of steps: double-double (error-free two-sum), so
// that the grouping of block sums into threads, warps and ranks does not show in the rounded result — together with
// the shard-independent segment length (pick_logL) this makes a sharded run return the bits of the single-GPU run
// (r01: Kahan here, and Ryser at n = 25 differed by 2.9e-11 between world sizes).
struct KahanC {
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    __device__ __forceinline__ void add(C128 x) { dd_add(re, x.re); dd_add(im, x.im); }
    __device__ __forceinline__ cdd get() const { cdd o; o.re = re; o.im = im; return o; }
};
struct KahanR {
    dd s = {0.0, 0.0};
    __device__ __forceinline__ void add(double x) { dd_add(s, x); }
    __device__ __forceinline__ cdd get() const { cdd o; o.re = s; o.im = {0.0, 0.0}; return o; }
};
struct AccI {
    unsigned long long s = 0;
    __device__ __forceinline__ void add(unsigned long long x) { s += x; }
};
template <typename S> struct AccOf;
template <> struct AccOf<C128> { using type = KahanC; };
template <> struct AccOf<double> { using type = KahanR; };
template <> struct AccOf<unsigned long long> { using type = AccI; };

// product of CPL values: WB_PERM_NC independent chains (default 4), or a balanced tree (WB_PERM_TREE)
#ifndef WB_PERM_NC
#define WB_PERM_NC 4
#endif
template <int CPL, typename S>
__device__ __forceinline__ S product(const S (&r)[CPL]) {
#ifdef WB_PERM_TREE
    S t[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) t[j] = r[j];
#pragma unroll
    for (int w = CPL; w > 1; w = (w + 1) / 2) {
#pragma unroll
        for (int j = 0; j < w / 2; ++j) t[j] = cmul(t[2 * j], t[2 * j + 1]);
        if (w & 1) t[w / 2] = t[w - 1];
    }
    return t[0];
#else
    constexpr int NC = CPL < WB_PERM_NC ? CPL : WB_PERM_NC;
    S p[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) p[j] = r[j];
#pragma unroll
    for (int c = NC; c < CPL; ++c) p[c % NC] = cmul(p[c % NC], r[c]);
#pragma unroll
    for (int w = NC; w > 1; w = (w + 1) / 2) {
#pragma unroll
        for (int j = 0; j < w / 2; ++j) p[j] = cmul(p[2 * j], p[2 * j + 1]);
        if (w & 1) p[w / 2] = p[w - 1];
    }
    return p[0];
#endif
}

// x, -x or 0 by selects (no FP64 issue slots)
__device__ __forceinline__ C128 weigh(C128 x, bool on, bool neg) {
    C128 y = {neg ? -x.re : x.re, neg ? -x.im : x.im};
    return {on ? y.re : 0.0, on ? y.im : 0.0};
}
__device__ __forceinline__ double weigh(double x, bool on, bool neg) { return on ? (neg ? -x : x) : 0.0; }
__device__ __forceinline__ unsigned long long weigh(unsigned long long x, bool on, bool neg) {
    return on ? (neg ? (0ull - x) : x) : 0ull;
}

// M in shared memory: sM[row * NP + col], NP = CPL * SL >= n, padded columns hold 0 and their sums are pinned
// to 1.  Lane `sub` of a lane group owns columns c * SL + sub (interleaved, so the SL pieces of one LDS.128 sit
// in adjacent 16-byte slots: no bank conflicts) of NS independent streams: a register tile r[NS][CPL].  Every
// row element fetched from shared memory feeds NS updates, which is what lifts the kernel off the shared-memory
// return path (one LDS.128 per 6 FP64 instructions saturates the 128 B/clk LSU before the FP64 pipe:
// tools/fp64_mix.cu, profiles/r01_fp64_mix.txt) and gives NS-fold instruction-level parallelism.
// All streams of the grid sit at the same offset t of their 2^logL-aligned segment, so the flipped row
// ctz(t) and the loop trip counts are uniform; steps outside [k0, k1) are masked by a zero weight.
template <int CPL, int SL, int NS, int MINB, typename S>
__global__ void __launch_bounds__(PERM_THREADS, MINB)
perm_kernel(const S* __restrict__ Mg, int n, int ryser, uint64_t k0, uint64_t k1, int logL,
            double* __restrict__ partials, unsigned long long* __restrict__ iout) {
    using Coef = typename Traits<S>::Coef;
    constexpr int NP = CPL * SL;
#ifndef WB_PERM_G1
#define WB_PERM_G1 2          // steps per group when one lane owns a stream (even)
#endif
    constexpr int G = SL < 2 ? WB_PERM_G1 : SL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S* sM = reinterpret_cast<S*>(smem_raw);
    for (int i = threadIdx.x; i < n * NP; i += PERM_THREADS) {
        const int row = i / NP, col = i % NP;
        sM[i] = (col < n) ? Mg[row * n + col] : Traits<S>::zero();
    }
    __syncthreads();

    typename AccOf<S>::type acc;
    const int sub = threadIdx.x & (SL - 1);
    const uint32_t L = 1u << logL;
    const uint64_t c_first = k0 >> logL, c_last = (k1 + L - 1) >> logL;  // chunk ids [c_first, c_last)
    const uint64_t ngroups = (uint64_t)gridDim.x * (PERM_THREADS / SL);
    const uint64_t group = ((uint64_t)blockIdx.x * PERM_THREADS + threadIdx.x) / SL;
    const Coef step = ryser ? (Coef)1 : (Coef)2;
    const S* mcol = sM + sub;
    const uint64_t per_round = ngroups * NS;
    const uint64_t rounds = (c_last - c_first + per_round - 1) / per_round;
    for (uint64_t it = 0; it < rounds; ++it) {
        uint64_t kbase[NS];
        uint32_t tlo[NS], thi[NS];          // steps t in [tlo, thi) of the segment lie in [k0, k1)
        S r[NS][CPL];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const uint64_t c = c_first + (it * ngroups + group) * NS + s;
            kbase[s] = c << logL;
            const uint64_t lo = kbase[s] < k0 ? k0 : kbase[s];
            uint64_t hi = kbase[s] + L;
            if (hi > k1) hi = k1;
            const bool live = c < c_last && lo < hi;
            tlo[s] = live ? (uint32_t)(lo - kbase[s]) : 0u;
            thi[s] = live ? (uint32_t)(hi - kbase[s]) : 0u;
#pragma unroll
            for (int q = 0; q < CPL; ++q) r[s][q] = Traits<S>::zero();
        }
        // ---- seed the column sums from g(kbase)
        for (int row = 0; row < n; ++row) {
            Coef d[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const bool set = ((kbase[s] ^ (kbase[s] >> 1)) >> row) & 1ull;
                d[s] = ryser ? (set ? (Coef)-1 : (Coef)0) : (set ? (Coef)-1 : (Coef)1);
            }
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const S mv = mcol[row * NP + q * SL];
#pragma unroll
                for (int s = 0; s < NS; ++s) axpy(r[s][q], d[s], mv);
            }
        }
#pragma unroll
        for (int q = 0; q < CPL; ++q)
            if (q * SL + sub >= n) {
#pragma unroll
                for (int s = 0; s < NS; ++s) r[s][q] = Traits<S>::one();
            }
        // ---- sweep in groups of G steps
        S blk[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) blk[s] = Traits<S>::zero();
        int inblk = 0;
        for (uint32_t t = 0; t < L; t += G) {
            S P[NS][G];
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const uint32_t tt = t + i;
                // flip the row that turns g(k - 1) into g(k), k = kbase + tt; nothing to do for the seeded step
                int row = tt ? (__ffs((int)tt) - 1) : 0;
                if (row >= n) row = n - 1;        // only in masked steps of a one-segment problem
                Coef d[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const uint32_t two = (uint32_t)((kbase[s] + tt) >> row) & 3u;   // bits row, row + 1 of k
                    const bool set = ((two ^ (two >> 1)) & 1u) != 0;                 // new value of Gray bit `row`
                    d[s] = tt ? (set ? -step : step) : (Coef)0;
                }
                const S* mrow = mcol + row * NP;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const S mv = mrow[q * SL];
#pragma unroll
                    for (int s = 0; s < NS; ++s) axpy(r[s][q], d[s], mv);          // padded columns add 0
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) P[s][i] = product<CPL, S>(r[s]);
            }
            // ---- transpose-reduce over the SL lanes: lane `sub` ends with the full product of step t + sub
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                S T;
                if constexpr (SL == 1) {
                    const bool on0 = t >= tlo[s] && t < thi[s], on1 = t + 1 >= tlo[s] && t + 1 < thi[s];
                    T = cadd(weigh(P[s][0], on0, false), weigh(P[s][1], on1, true));     // t is even: + then -
#pragma unroll
                    for (int i = 2; i < G; ++i)
                        T = cadd(T, weigh(P[s][i], t + i >= tlo[s] && t + i < thi[s], (i & 1) != 0));
                } else if constexpr (SL == 2) {
                    const S mine = sub ? P[s][1] : P[s][0], give = sub ? P[s][0] : P[s][1];
                    const S got = shfl_xor_s(0xffffffffu, give, 1);
                    const uint32_t tt = t + sub;
                    T = weigh(cmul(mine, got), tt >= tlo[s] && tt < thi[s], sub & 1);
                } else {
                    const bool hi = sub & 2;
                    const S keep0 = hi ? P[s][2] : P[s][0], keep1 = hi ? P[s][3] : P[s][1];
                    const S give0 = hi ? P[s][0] : P[s][2], give1 = hi ? P[s][1] : P[s][3];
                    const S q0 = cmul(keep0, shfl_xor_s(0xffffffffu, give0, 2));
                    const S q1 = cmul(keep1, shfl_xor_s(0xffffffffu, give1, 2));
                    const bool odd = sub & 1;
                    const S mine = odd ? q1 : q0, give = odd ? q0 : q1;
                    const uint32_t tt = t + sub;
                    T = weigh(cmul(mine, shfl_xor_s(0xffffffffu, give, 1)), tt >= tlo[s] && tt < thi[s], odd);
                }
                blk[s] = cadd(blk[s], T);
            }
            if (++inblk == PERM_BLOCK_GROUPS) {
#pragma unroll
                for (int s = 0; s < NS; ++s) { acc.add(blk[s]); blk[s] = Traits<S>::zero(); }
                inblk = 0;
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) acc.add(blk[s]);
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

constexpr int PERM_MAX_GRID = 4096;
constexpr int PERM_MAX_N = 64;

// Segment length 2^logL: long enough to amortise seeding (n / 6 steps' worth), short enough to fill the grid.  It is
// chosen from the FULL step count of the problem and the full grid, as if the range were split over 8 ranks — never
// from the shard [k0, k1) — because the rounding of the incrementally updated column sums depends on where a segment
// was seeded: with one segment length for every world size <= 8 each step's term is bit-identical wherever it runs.
static int pick_logL_for(uint64_t total, uint64_t nstreams, uint64_t segs_per_slot) {
    int logL = 6;
    while (logL < 20 && (total >> (logL + 1)) >= nstreams * segs_per_slot) ++logL;
    return logL;
}
static int pick_logL(uint64_t full_steps, uint64_t nstreams) {
    const uint64_t shard = full_steps >> 3;
    int logL = pick_logL_for(shard, nstreams, 32);            // >= 32 segments per stream slot at 8 ranks: <= 3 % tail
    if (logL < 9) {                                            // mid sizes: accept >= 8 segments per slot at 8 ranks rather
        const int alt = pick_logL_for(shard, nstreams, 8);     // than pay the seeding of 64-step segments on one GPU
        logL = alt < 9 ? alt : 9;
    }
    return logL;
}

template <int CPL, int SL, int NS, int MINB, typename S>
static int launch_perm(const S* dM, int n, int method, uint64_t k0, uint64_t k1, double* partials,
                       unsigned long long* iout, int* grid_out, cudaStream_t st) {
    int dev = 0, sms = 0, occ = 1;
    WB_CUDA(cudaGetDevice(&dev));
    if (device_sm_count(dev, &sms)) return WB200_ECUDA;
    constexpr int NP = CPL * SL;
    const size_t shm = sizeof(S) * (size_t)n * NP;
    WB_CUDA(cudaFuncSetAttribute(perm_kernel<CPL, SL, NS, MINB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, perm_kernel<CPL, SL, NS, MINB, S>, PERM_THREADS, shm));
    if (occ < 1) occ = 1;
    const uint64_t total = k1 - k0;
    const uint64_t spb = (PERM_THREADS / SL) * NS;                // streams per block
    uint64_t maxgrid = (uint64_t)sms * occ;
    if (maxgrid > PERM_MAX_GRID) maxgrid = PERM_MAX_GRID;
    uint64_t want = (total + spb * 64 - 1) / (spb * 64);
    int grid = (int)(want < maxgrid ? (want ? want : 1) : maxgrid);
    const int bits = method ? n : n - 1;          // the step index has `bits` bits: rows ctz(t) stay below n
    int logL = pick_logL(1ull << bits, maxgrid * spb);
    if (logL > bits) logL = bits;
    if (logL < 1) logL = 1;
    perm_kernel<CPL, SL, NS, MINB, S><<<grid, PERM_THREADS, shm, st>>>(dM, n, method, k0, k1, logL, partials, iout);
    WB_CUDA(cudaGetLastError());
    *grid_out = grid;
    return WB200_OK;
}

// Tile shape per matrix size (measured on B200, profiles/r01_perm_shapes.txt): one lane per stream with all n
// column sums in registers is fastest up to n = 40 (6.7e10 steps/s at n = 32, 5.5e10 at n = 40); beyond that the
// sums no longer fit one thread and four lanes share two streams.  WB200_PERM_SL / WB200_PERM_NS override.
static int env_int(const char* name) {
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

template <typename S>
static int dispatch_perm(const S* dM, int n, int method, uint64_t k0, uint64_t k1, double* partials,
                         unsigned long long* iout, int* grid, cudaStream_t st) {
    static const int f_sl = env_int("WB200_PERM_SL"), f_ns = env_int("WB200_PERM_NS");
    int sl = n <= 40 ? 1 : 4, ns = n <= 40 ? 1 : 2;
    if (f_sl == 2 && n > 8 && n <= 32) { sl = 2; ns = 2; }
    if (f_sl == 4 && n > 8) { sl = 4; ns = 2; }
    if ((f_ns == 1 || f_ns == 2) && sl > 1) ns = f_ns;
    const int cpl = (n + sl - 1) / sl;
#define WB_PERM_CASE(C, L, N) \
    if (sl == L && ns == N && cpl <= C) return launch_perm<C, L, N, PERM_MIN_CTAS, S>(dM, n, method, k0, k1, partials, iout, grid, st);
    WB_PERM_CASE(4, 1, 1) WB_PERM_CASE(8, 1, 1) WB_PERM_CASE(12, 1, 1) WB_PERM_CASE(16, 1, 1) WB_PERM_CASE(20, 1, 1)
    WB_PERM_CASE(24, 1, 1) WB_PERM_CASE(28, 1, 1) WB_PERM_CASE(30, 1, 1) WB_PERM_CASE(32, 1, 1) WB_PERM_CASE(34, 1, 1)
    WB_PERM_CASE(36, 1, 1) WB_PERM_CASE(38, 1, 1) WB_PERM_CASE(40, 1, 1)
    WB_PERM_CASE(8, 2, 1) WB_PERM_CASE(16, 2, 1) WB_PERM_CASE(8, 2, 2) WB_PERM_CASE(12, 2, 2) WB_PERM_CASE(16, 2, 2)
    WB_PERM_CASE(8, 4, 1) WB_PERM_CASE(12, 4, 1) WB_PERM_CASE(16, 4, 1)
    WB_PERM_CASE(4, 4, 2) WB_PERM_CASE(8, 4, 2) WB_PERM_CASE(10, 4, 2) WB_PERM_CASE(11, 4, 2) WB_PERM_CASE(12, 4, 2)
    WB_PERM_CASE(14, 4, 2) WB_PERM_CASE(16, 4, 2)
#undef WB_PERM_CASE
    set_error("perm: no kernel for n = %d with SL = %d, NS = %d (limit n <= %d)", n, sl, ns, PERM_MAX_N);
    return WB200_ENOSUP;
}

int check_perm_args(int n, int method, uint64_t k0, uint64_t k1) {
    if (n < 1 || n > PERM_MAX_N) {
        set_error("perm: n = %d outside [1, %d]", n, PERM_MAX_N);
        return n > PERM_MAX_N ? WB200_ENOSUP : WB200_EINVAL;
    }
    if (method != 0 && method != 1) { set_error("perm: method must be 0 (bbfg) or 1 (ryser)"); return WB200_EINVAL; }
    const int bits = method ? n : n - 1;
    if (bits > 62) { set_error("perm: 2^%d steps exceed the 64-bit step index", bits); return WB200_ENOSUP; }
    const uint64_t steps = 1ull << bits;
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


// =================================================================================================
// Bristolian (thewalrus/_permanent.py:198-249): sum over row subsets Y of the m x n matrix A of
// (-1)^(m - |Y|) perm(A_Y^H A_Y + E), each permanent by the Glynn/BBFG Gray-code sum above.
// =================================================================================================
// One warp per work unit = (outer subset j, inner chunk): the warp builds the n x n Gram matrix of its subset in
// its slice of shared memory, its 32 lanes sweep aligned equal segments of the inner Gray code (same flipped row
// in every lane), and lane 0 adds sign * (warp sum) to the warp's double-double total.  Units are dealt to warps
// round-robin; partial totals are combined in fixed order.
constexpr int BRS_WARPS = 8;

struct BrsParams {
    const C128* A;     // m x n
    const C128* E;     // n x n or null
    int m, n, log_chunks, log_seg;   // inner steps 2^(n-1) = chunks * 32 lanes * 2^log_seg (or fewer lanes, see kernel)
    unsigned long long j0, j1;
    double* partials;  // (gridDim.x * BRS_WARPS) x 4
};

template <int NP>
__global__ void __launch_bounds__(32 * BRS_WARPS) brs_kernel(BrsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n, m = p.m;
    C128* G = reinterpret_cast<C128*>(smem_raw) + (size_t)warp * n * NP;     // G[row * NP + col]
    const unsigned long long gw = (unsigned long long)blockIdx.x * BRS_WARPS + warp;
    const unsigned long long nw = (unsigned long long)gridDim.x * BRS_WARPS;
    const unsigned long long units = (p.j1 - p.j0) << p.log_chunks;
    const unsigned long long inner = 1ull << (n - 1);
    const unsigned long long seg = 1ull << p.log_seg;                         // steps per lane per unit
    cdd tot;
    tot.re = {0.0, 0.0}; tot.im = {0.0, 0.0};
    for (unsigned long long u = gw; u < units; u += nw) {
        const unsigned long long j = p.j0 + (u >> p.log_chunks), chunk = u & ((1ull << p.log_chunks) - 1ull);
        // ---- Gram matrix of the kept rows (bit i of the MSB-first label keeps row i: find_kept_edges)
        __syncwarp();
        for (int idx = lane; idx < n * NP; idx += 32) {
            const int a = idx / NP, b = idx - a * NP;
            C128 g = {0.0, 0.0};
            if (b < n) {
                if (p.E) g = p.E[a * n + b];
                for (int y = 0; y < m; ++y) {
                    if ((j >> (m - 1 - y)) & 1ull) {
                        const C128 ya = p.A[y * n + a], yb = p.A[y * n + b];
                        g.re += ya.re * yb.re + ya.im * yb.im;        // conj(ya) * yb
                        g.im += ya.re * yb.im - ya.im * yb.re;
                    }
                }
            }
            G[idx] = g;
        }
        __syncwarp();
        const int kept = __popcll(j);
        // ---- this lane's segment of the inner Gray code
        const unsigned long long kb = (chunk * 32ull + (unsigned long long)lane) << p.log_seg;
        KahanC acc;
        if (kb < inner) {
            const unsigned long long gray = kb ^ (kb >> 1);
            C128 r[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) r[q] = {0.0, 0.0};
            for (int row = 0; row < n; ++row) {
                const double d = ((gray >> row) & 1ull) ? -1.0 : 1.0;
#pragma unroll
                for (int q = 0; q < NP; ++q) axpy(r[q], d, G[row * NP + q]);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q)
                if (q >= n) r[q] = {1.0, 0.0};
            acc.add(weigh(product<NP, C128>(r), true, (kb & 1ull) != 0));
            for (unsigned long long t = 1; t < seg; ++t) {
                const unsigned long long k = kb + t;
                const int row = __ffsll((long long)t) - 1;            // = ctz(k): kb is a multiple of seg
                const bool set = ((k ^ (k >> 1)) >> row) & 1ull;
                const double d = set ? -2.0 : 2.0;
                const C128* grow = G + row * NP;
#pragma unroll
                for (int q = 0; q < NP; ++q) axpy(r[q], d, grow[q]);
                acc.add(weigh(product<NP, C128>(r), true, (k & 1ull) != 0));
            }
        }
        cdd v = acc.get();
        v.re = warp_reduce_dd(v.re);
        v.im = warp_reduce_dd(v.im);
        if (lane == 0) {
            const double sg = ((m - kept) & 1) ? -1.0 : 1.0;
            dd_add_dd(tot.re, dd{sg * v.re.hi, sg * v.re.lo});
            dd_add_dd(tot.im, dd{sg * v.im.hi, sg * v.im.lo});
        }
    }
    if (lane == 0) {
        double* o = p.partials + gw * 4;
        o[0] = tot.re.hi; o[1] = tot.re.lo; o[2] = tot.im.hi; o[3] = tot.im.lo;
    }
}

template <int NP>
static int launch_brs(BrsParams p, int grid, cudaStream_t st) {
    const size_t shm = sizeof(C128) * (size_t)p.n * NP * BRS_WARPS;
    WB_CUDA(cudaFuncSetAttribute(brs_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    brs_kernel<NP><<<grid, 32 * BRS_WARPS, shm, st>>>(p);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

constexpr int BRS_MAX_N = 32, BRS_MAX_M = 40;

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

extern "C" int wb200_brs_dev(const double* dA, const double* dE, int m, int n, uint64_t j0, uint64_t j1, double* d_out4,
                             void* stream) {
    if (!dA || !d_out4) { set_error("brs: null pointer"); return WB200_EINVAL; }
    if (m < 1 || n < 1) { set_error("brs: A must be m x n with m, n >= 1 (got %d x %d)", m, n); return WB200_EINVAL; }
    if (n > BRS_MAX_N || m > BRS_MAX_M) { set_error("brs: %d x %d exceeds the kernel limits (%d rows, %d columns)", m, n, BRS_MAX_M, BRS_MAX_N); return WB200_ENOSUP; }
    const uint64_t outer = 1ull << m;
    if (j0 > j1 || j1 > outer) { set_error("brs: bad subset range"); return WB200_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0, sms = 0;
    (void)cudaGetLastError();
    WB_CUDA(cudaGetDevice(&device));
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    BrsParams p;
    p.A = (const C128*)dA; p.E = (const C128*)dE; p.m = m; p.n = n; p.j0 = j0; p.j1 = j1;
    // inner Gray code of 2^(n-1) steps = chunks x 32 lanes x 2^log_seg; more chunks when there are few outer subsets
    const int inner_bits = n - 1;
    int log_seg = inner_bits > 5 ? inner_bits - 5 : 0, log_chunks = 0;
    const uint64_t nouter = j1 - j0, target = (uint64_t)sms * BRS_WARPS * 4;
    while (log_seg > 10 && (nouter << log_chunks) < target) { --log_seg; ++log_chunks; }
    p.log_seg = log_seg; p.log_chunks = log_chunks;
    const uint64_t units = nouter << log_chunks;
    int grid = (int)((units + BRS_WARPS - 1) / BRS_WARPS);
    if (grid > sms * 2) grid = sms * 2;
    if (grid < 1) grid = 1;
    StreamBuf dpart;
    WB_POOL(dpart.alloc(sizeof(double) * 4 * grid * BRS_WARPS, st));
    WB_CUDA(cudaMemsetAsync(dpart.p, 0, sizeof(double) * 4 * grid * BRS_WARPS, st));
    p.partials = (double*)dpart.p;
    int rc;
    if (n <= 4) rc = launch_brs<4>(p, grid, st);
    else if (n <= 8) rc = launch_brs<8>(p, grid, st);
    else if (n <= 12) rc = launch_brs<12>(p, grid, st);
    else if (n <= 16) rc = launch_brs<16>(p, grid, st);
    else if (n <= 20) rc = launch_brs<20>(p, grid, st);
    else if (n <= 24) rc = launch_brs<24>(p, grid, st);
    else if (n <= 28) rc = launch_brs<28>(p, grid, st);
    else rc = launch_brs<32>(p, grid, st);
    if (rc) return rc;
    final_reduce_kernel<<<1, 32, 0, st>>>((const double*)dpart.p, grid * BRS_WARPS, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

extern "C" int wb200_brs_host(int device, const double* A, const double* E, int m, int n, uint64_t j0, uint64_t j1,
                              double out4[4], double* kernel_ms) {
    if (!A || !out4) { set_error("brs: null pointer"); return WB200_EINVAL; }
    if (m < 1 || n < 1) { set_error("brs: A must be m x n with m, n >= 1 (got %d x %d)", m, n); return WB200_EINVAL; }
    if (n > BRS_MAX_N || m > BRS_MAX_M) { set_error("brs: %d x %d exceeds the kernel limits (%d rows, %d columns)", m, n, BRS_MAX_M, BRS_MAX_N); return WB200_ENOSUP; }
    WB_CUDA(cudaSetDevice(device));
    struct Buf { void* p = nullptr; ~Buf() { if (p) pool_free(p); } } dA, dE, dout;
    WB_POOL(pool_alloc(&dA.p, sizeof(C128) * m * n));
    WB_CUDA(cudaMemcpy(dA.p, A, sizeof(C128) * m * n, cudaMemcpyHostToDevice));
    if (E) {
        WB_POOL(pool_alloc(&dE.p, sizeof(C128) * n * n));
        WB_CUDA(cudaMemcpy(dE.p, E, sizeof(C128) * n * n, cudaMemcpyHostToDevice));
    }
    WB_POOL(pool_alloc(&dout.p, sizeof(double) * 4));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) {
        WB_CUDA(cudaEventCreate(&e0));
        WB_CUDA(cudaEventCreate(&e1));
        WB_CUDA(cudaEventRecord(e0, 0));
    }
    int rc = wb200_brs_dev((const double*)dA.p, (const double*)dE.p, m, n, j0, j1, (double*)dout.p, nullptr);
    if (rc) { if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); } return rc; }
    if (kernel_ms) {
        WB_CUDA(cudaEventRecord(e1, 0));
        WB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        WB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *kernel_ms = ms;
    }
    WB_CUDA(cudaMemcpy(out4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return WB200_OK;
}
