// FP64 instruction-mix micro-benchmark for B200 (sm_100a): which rates do DFMA / DMUL / DADD and the
// Glynn-permanent inner loop (complex update + complex product chain) reach, at which occupancy?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix tools/fp64_mix.cu
// Rates are printed as FP64 warp-instructions per clock per SM sub-partition; the pipe's peak is 0.5
// (16 lanes per sub-partition: 64 FMA/clk/SM).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct C { double re, im; };
__device__ __forceinline__ C cmul(C a, C b) { return {fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)}; }
// all-FMA complex multiply (no DMUL): the first product of each component is an FMA onto -0.0
__device__ __forceinline__ C cmul_fma(C a, C b, double nz) {
    return {fma(a.re, b.re, fma(-a.im, b.im, nz)), fma(a.re, b.im, fma(a.im, b.re, nz))};
}

template <int OP, int CH>
__global__ void __launch_bounds__(256) k_basic(double* out, int iters, double a, double b) {
    double x[CH], y[CH], z[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = a + i * 1e-9; z[i] = b + i * 1e-12; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (OP == 0) x[i] = fma(x[i], y[i], z[i]);
                if (OP == 1) x[i] = x[i] * y[i];
                if (OP == 2) x[i] = x[i] + z[i];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NCH independent complex product chains
template <int NCH, int FMAONLY>
__global__ void __launch_bounds__(256) k_cmul(double* out, int iters, double a, double b, double nz) {
    C p[NCH], w[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { p[i] = {1.0 + threadIdx.x * 1e-6, 1e-3 * i}; w[i] = {a, b + 1e-9 * i}; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) p[i] = FMAONLY ? cmul_fma(p[i], w[i], nz) : cmul(p[i], w[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += p[i].re + p[i].im;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// permanent-like step: r += d * m (CPL complex, m in registers or shared memory), P = prod r, acc += P
template <int CPL, int LDS, int FMAONLY>
__global__ void __launch_bounds__(256) k_perm(double* out, int iters, double a, double b, double nz) {
    __shared__ C tab[64 * CPL];
    for (int i = threadIdx.x; i < 64 * CPL; i += blockDim.x) tab[i] = {1e-3 * (i % 7) - 3e-3, 1e-3 * (i % 5) - 2e-3};
    __syncthreads();
    C r[CPL], m[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) { r[i] = {1.0 + threadIdx.x * 1e-6, 1e-3 * i}; m[i] = {a * 1e-3, b * 1e-3 + 1e-9 * i}; }
    C acc = {0, 0};
    for (int it = 0; it < iters; ++it) {
        const double d = (it & 1) ? 2.0 : -2.0;
        const C* row = tab + (it & 63) * CPL;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const C mv = LDS ? row[i] : m[i];
            r[i].re = fma(d, mv.re, r[i].re);
            r[i].im = fma(d, mv.im, r[i].im);
        }
        C p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = r[j];
#pragma unroll
        for (int c = 4; c < CPL; ++c) p[c & 3] = FMAONLY ? cmul_fma(p[c & 3], r[c], nz) : cmul(p[c & 3], r[c]);
        C q0 = FMAONLY ? cmul_fma(p[0], p[1], nz) : cmul(p[0], p[1]);
        C q1 = FMAONLY ? cmul_fma(p[2], p[3], nz) : cmul(p[2], p[3]);
        C q = FMAONLY ? cmul_fma(q0, q1, nz) : cmul(q0, q1);
        acc.re += q.re; acc.im += q.im;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.re + acc.im;
}

// permanent-like step with the matrix in the kernel-parameter constant bank and a warp-uniform row index:
// the row values become uniform-register operands of the update DFMAs (no LDS write-back traffic)
struct __align__(16) C16 { double re, im; };
template <int NP> struct Mat { C16 m[NP * NP]; };
template <int NP, int MODE>
__global__ void __launch_bounds__(128) k_permc(const __grid_constant__ Mat<NP> M, double* out, int logL, unsigned long long kb0) {
    C r[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) r[q] = {1.0 + threadIdx.x * 1e-6, 1e-3 * q};
    const unsigned long long kb = kb0 + ((unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) << logL);
    C acc = {0, 0};
    const int L = 1 << logL;
    for (int t = 1; t < L; ++t) {
        const int row = MODE == 1 ? 0 : (__ffs(t) - 1) % NP;                 // uniform
        const unsigned long long k = kb + t;
        const bool set = ((k ^ (k >> 1)) >> row) & 1;
        const double d = set ? -2.0 : 2.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const C16 mv = M.m[row * NP + q];
            r[q].re = fma(d, mv.re, r[q].re);
            r[q].im = fma(d, mv.im, r[q].im);
        }
        if (MODE == 2) { acc.re += r[t & 7].re; continue; }   // updates only
        C p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = r[j];
#pragma unroll
        for (int c = 4; c < NP; ++c) p[c & 3] = cmul(p[c & 3], r[c]);
        C q = cmul(cmul(p[0], p[1]), cmul(p[2], p[3]));
        acc.re += q.re; acc.im += q.im;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.re + acc.im;
}

// NS streams per thread sharing every row element loaded from shared memory.  MODE 0: update + product,
// 1: update only, 2: product only (r perturbed by one DADD per stream so the chain is not hoisted)
template <int CPL, int NS, int MODE>
__global__ void __launch_bounds__(128, 2) k_tile(double* out, int iters, double a, double b) {
    __shared__ C16 tab[64 * CPL];
    for (int i = threadIdx.x; i < 64 * CPL; i += blockDim.x) tab[i] = {1e-3 * (i % 7) - 3e-3, 1e-3 * (i % 5) - 2e-3};
    __syncthreads();
    C r[NS][CPL];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < CPL; ++i) r[s][i] = {1.0 + threadIdx.x * 1e-6 + s, 1e-3 * i};
    C acc = {0, 0};
    for (int it = 0; it < iters; ++it) {
        double d[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) d[s] = ((it + s + threadIdx.x) & 1) ? 2.0 : -2.0;
        const C16* row = tab + (it & 63) * CPL;
        if (MODE != 2) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const C16 mv = row[i];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    r[s][i].re = fma(d[s], mv.re, r[s][i].re);
                    r[s][i].im = fma(d[s], mv.im, r[s][i].im);
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < NS; ++s) r[s][it & (CPL - 1)].re += d[s];
        }
        if (MODE == 1) { acc.re += r[0][it & (CPL - 1)].re; continue; }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            C p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = r[s][j];
#pragma unroll
            for (int c = 4; c < CPL; ++c) p[c & 3] = cmul(p[c & 3], r[s][c]);
            C q = cmul(cmul(p[0], p[1]), cmul(p[2], p[3]));
            acc.re += q.re; acc.im += q.im;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.re + acc.im;
}

template <typename F>
static double timeit(F launch, int reps = 3) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best * 1e-3;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const double clk = p.clockRate * 1e3;
    printf("device %s sms %d clock %.0f MHz\n", p.name, sms, clk * 1e-6);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 256));
    const int iters = 4000;
    const double nz = -0.0;
    auto rate = [&](double t, double warp_instr_per_warp, int grid, int threads) {
        const double total = warp_instr_per_warp * grid * (threads / 32);
        return total / (t * clk) / (sms * 4.0);
    };
    for (int wps = 1; wps <= 2; ++wps) {
        const int threads = 128, grid = sms * wps, it8 = iters * 8;
        double t;
#define RUNT(name, kern, instr) t = timeit([&] { kern; }); printf("%-30s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", name, wps, t * 1e3, rate(t, instr, grid, threads));
        RUNT("tile CPL16 NS2 upd+prod", (k_tile<16, 2, 0><<<grid, threads>>>(out, it8, 1.0, 1e-4)), 2 * (32.0 + 60 + 2) * it8)
        RUNT("tile CPL16 NS2 upd only", (k_tile<16, 2, 1><<<grid, threads>>>(out, it8, 1.0, 1e-4)), (2 * 32.0 + 1) * it8)
        RUNT("tile CPL16 NS2 prod only", (k_tile<16, 2, 2><<<grid, threads>>>(out, it8, 1.0, 1e-4)), 2 * (60.0 + 2 + 1) * it8)
        RUNT("tile CPL16 NS1 upd+prod", (k_tile<16, 1, 0><<<grid, threads>>>(out, it8, 1.0, 1e-4)), 1 * (32.0 + 60 + 2) * it8)
        RUNT("tile CPL16 NS1 upd only", (k_tile<16, 1, 1><<<grid, threads>>>(out, it8, 1.0, 1e-4)), (1 * 32.0 + 1) * it8)
        RUNT("tile CPL16 NS1 prod only", (k_tile<16, 1, 2><<<grid, threads>>>(out, it8, 1.0, 1e-4)), 1 * (60.0 + 2 + 1) * it8)
        RUNT("tile CPL8 NS4 upd+prod", (k_tile<8, 4, 0><<<grid, threads>>>(out, it8, 1.0, 1e-4)), 4 * (16.0 + 28 + 2) * it8)
        RUNT("tile CPL8 NS4 upd only", (k_tile<8, 4, 1><<<grid, threads>>>(out, it8, 1.0, 1e-4)), (4 * 16.0 + 1) * it8)
        RUNT("tile CPL8 NS4 prod only", (k_tile<8, 4, 2><<<grid, threads>>>(out, it8, 1.0, 1e-4)), 4 * (28.0 + 2 + 1) * it8)
    }
    if (0) {
        static Mat<32> M32; static Mat<40> M40; static Mat<16> M16;
        for (int i = 0; i < 32 * 32; ++i) M32.m[i] = {1e-3 * (i % 7) - 3e-3, 1e-3 * (i % 5) - 2e-3};
        for (int i = 0; i < 40 * 40; ++i) M40.m[i] = {1e-3 * (i % 7) - 3e-3, 1e-3 * (i % 5) - 2e-3};
        for (int i = 0; i < 16 * 16; ++i) M16.m[i] = {1e-3 * (i % 7) - 3e-3, 1e-3 * (i % 5) - 2e-3};
        const int logL = 14;
        for (int wps = 1; wps <= 4; ++wps) {
            const int threads = 128, grid = sms * wps;
            double t;
            t = timeit([&] { k_permc<16, 0><<<grid, threads>>>(M16, out, logL, 12345ull << 20); });
            printf("%-26s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", "permc NP16 (UR operands)", wps, t * 1e3, rate(t, (32.0 + 60 + 2) * ((1 << logL) - 1), grid, threads));
            if (wps <= 3) {
            t = timeit([&] { k_permc<32, 0><<<grid, threads>>>(M32, out, logL, 12345ull << 20); });
            printf("%-26s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", "permc NP32 (UR operands)", wps, t * 1e3, rate(t, (64.0 + 124 + 2) * ((1 << logL) - 1), grid, threads));
            }
            if (wps <= 3) {
            t = timeit([&] { k_permc<32, 1><<<grid, threads>>>(M32, out, logL, 12345ull << 20); });
            printf("%-26s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", "permc NP32 static row", wps, t * 1e3, rate(t, (64.0 + 124 + 2) * ((1 << logL) - 1), grid, threads));
            t = timeit([&] { k_permc<32, 2><<<grid, threads>>>(M32, out, logL, 12345ull << 20); });
            printf("%-26s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", "permc NP32 updates only", wps, t * 1e3, rate(t, (64.0 + 1) * ((1 << logL) - 1), grid, threads));
            }
            if (wps <= 2) {
            t = timeit([&] { k_permc<40, 0><<<grid, threads>>>(M40, out, logL, 12345ull << 20); });
            printf("%-26s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", "permc NP40 (UR operands)", wps, t * 1e3, rate(t, (80.0 + 156 + 2) * ((1 << logL) - 1), grid, threads));
            }
        }
    }
    for (int wps = 1; wps <= 0; wps *= 2) {   // warps per sub-partition (disabled in this run)
        const int threads = 128, grid = sms * wps;
        double t;
#define RUN(name, kern, instr) t = timeit([&] { kern; }); printf("%-26s warps/SMSP %d : %8.3f ms  %.3f fp64 instr/clk/SMSP (peak 0.5)\n", name, wps, t * 1e3, rate(t, instr, grid, threads));
        RUN("dfma distinct ch8", (k_basic<0, 8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9)), 64.0 * iters)
        RUN("dmul ch8", (k_basic<1, 8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9)), 64.0 * iters)
        RUN("dadd ch8", (k_basic<2, 8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9)), 64.0 * iters)
        RUN("cmul chains x4", (k_cmul<4, 0><<<grid, threads>>>(out, iters, 1.0000001, 1e-4, nz)), 4.0 * 4 * 8 * iters)
        RUN("cmul chains x8", (k_cmul<8, 0><<<grid, threads>>>(out, iters, 1.0000001, 1e-4, nz)), 4.0 * 8 * 8 * iters)
        RUN("cmul chains x8 fma-only", (k_cmul<8, 1><<<grid, threads>>>(out, iters, 1.0000001, 1e-4, nz)), 4.0 * 8 * 8 * iters)
        RUN("perm step CPL8 regs", (k_perm<8, 0, 0><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-4, nz)), (16.0 + 28 + 2) * iters * 8)
        RUN("perm step CPL16 regs", (k_perm<16, 0, 0><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-4, nz)), (32.0 + 60 + 2) * iters * 8)
        RUN("perm step CPL8 lds", (k_perm<8, 1, 0><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-4, nz)), (16.0 + 28 + 2) * iters * 8)
        RUN("perm step CPL16 lds", (k_perm<16, 1, 0><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-4, nz)), (32.0 + 60 + 2) * iters * 8)
        RUN("perm step CPL16 lds fma", (k_perm<16, 1, 1><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-4, nz)), (32.0 + 60 + 2) * iters * 8)
        RUN("perm step CPL32 lds", (k_perm<32, 1, 0><<<grid, threads>>>(out, iters * 8, 1.0000001, 1e-4, nz)), (64.0 + 124 + 2) * iters * 8)
    }
    return 0;
}
