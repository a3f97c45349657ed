// General loop-hafnian subset sum: repeated edges (mixed-radix index), Glynn or inclusion/exclusion,
// optional loops (D) and optional unpaired vertex (oddloop / oddV).
//
// Replaces the prange bodies of _calc_hafnian / _calc_loop_hafnian for arbitrary edge_reps
// (thewalrus/_hafnian.py:416-467, 512-577): find_kept_edges (:162-180), binomial weights (:447-449),
// get_submatrices (:315-356: rows/cols with delta = 0 deleted, halves swapped, columns scaled by delta),
// f / f_loop / f_loop_odd (:183-285).
//
// One CTA per subset (grid-stride).  The reduced matrix M = AX_S lives in shared memory; power traces
// tr(M^k) come from a product chain with the pairing tr(M^(a+b)) = sum_rc (M^a)[r,c] (M^b)[c,r], continued
// past the matrix size where the reference switches to La Budde + Newton (thewalrus/charpoly.py:319-326) —
// the same numbers, without the Hessenberg reduction.  The series coefficients use
// c_t = (1/t) sum_i i a_i c_(t-i), the closed recurrence of the reference's `comb` loops.
#include "common.cuh"

namespace wb {

constexpr int LG_THREADS = 128;
constexpr int LG_MAX_EDGES = 32;
constexpr int LG_MAX_ORDER = 160;  // max series order N (odd case) / N/2 (even case)

struct LgParams {
    const double2* A;   // n x n
    const double2* D;   // n or null
    const double2* oddV;  // n or null
    double2 oddloop;
    int n, E, glynn, has_odd, N;  // N = total photon number
    int reps[LG_MAX_EDGES];
    uint64_t j0, j1;
};

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ void cfma2(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ double2 block_sum(double2 v, double2* scratch) {
    // all LG_THREADS threads call; result valid in every thread
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v.x += shfl_xor_d(v.x, off);
        v.y += shfl_xor_d(v.y, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double2 t = scratch[0];
#pragma unroll
    for (int w = 1; w < LG_THREADS / 32; ++w) { t.x += scratch[w].x; t.y += scratch[w].y; }
    return t;
}

__global__ void __launch_bounds__(LG_THREADS) lhaf_general_kernel(LgParams p, double* __restrict__ partials) {
    extern __shared__ __align__(16) double smem_lg[];
    const int n = p.n, E = p.E;
    double2* M = reinterpret_cast<double2*>(smem_lg);  // s x s (stride s)
    double2* P = M + n * n;
    double2* Pn = P + n * n;
    double2* vXD = Pn + n * n;     // n
    double2* vD = vXD + n;         // n  (advanced: M^t D)
    double2* vD2 = vD + n;         // n  scratch
    double2* vOV = vD2 + n;        // n  oddVX
    double2* ptr = vOV + n;        // LG_MAX_ORDER + 2 power traces
    double2* fac = ptr + LG_MAX_ORDER + 2;   // series factors i * a_i
    double2* cser = fac + LG_MAX_ORDER + 2;  // series coefficients
    double2* scratch = cser + LG_MAX_ORDER + 2;  // 8
    __shared__ int s_rows[2 * LG_MAX_EDGES];
    __shared__ double s_delta[2 * LG_MAX_EDGES];
    __shared__ int s_k, s_esum, s_d0zero;
    __shared__ double s_weight;

    const int tid = threadIdx.x;
    const bool loops = p.D != nullptr;
    const int N = p.N;
    const int tr_order = N / 2;                   // power traces p_1..p_{N/2}
    const int order = p.has_odd ? N : N / 2;      // series order
    cdd acc;
    acc.re = {0.0, 0.0};
    acc.im = {0.0, 0.0};

    for (uint64_t j = p.j0 + blockIdx.x; j < p.j1; j += gridDim.x) {
        __syncthreads();
        if (tid == 0) {
            // mixed-radix digits, most significant first (find_kept_edges)
            uint64_t num = j;
            int kept[LG_MAX_EDGES];
            for (int i = E - 1; i >= 0; --i) {
                const uint64_t base = (uint64_t)p.reps[i] + 1;
                kept[i] = (int)(num % base);
                num /= base;
            }
            int k = 0, esum = 0;
            double w = 1.0;
            for (int i = 0; i < E; ++i) {
                esum += kept[i];
                // binomial C(reps, kept) in floating point, exact for the sizes involved
                double b = 1.0;
                const int r = p.reps[i], kk = kept[i] < r - kept[i] ? kept[i] : r - kept[i];
                for (int q = 0; q < kk; ++q) b = b * (double)(r - q) / (double)(q + 1);
                w *= rint(b);
                const int d = p.glynn ? 2 * kept[i] - r : kept[i];
                if (i == 0) s_d0zero = (d == 0);
                if (d != 0) { s_rows[k] = i; s_delta[k] = (double)d; ++k; }
            }
            for (int a = 0; a < k; ++a) { s_rows[k + a] = s_rows[a] + E; s_delta[k + a] = s_delta[a]; }
            s_k = k; s_esum = esum; s_weight = w;
        }
        __syncthreads();
        const int k = s_k, s = 2 * k;
        // ---- M[r][c] = delta_c * A[row r, row sigma(c)]; vectors
        for (int idx = tid; idx < s * s; idx += LG_THREADS) {
            const int r = idx / s, c = idx % s;
            const int sc = c < k ? c + k : c - k;
            double2 a = p.A[(size_t)s_rows[r] * n + s_rows[sc]];
            const double d = s_delta[c];
            M[idx] = make_double2(a.x * d, a.y * d);
            P[idx] = M[idx];
        }
        for (int c = tid; c < s; c += LG_THREADS) {
            const int sc = c < k ? c + k : c - k;
            if (loops) {
                double2 dv = p.D[s_rows[sc]];
                vXD[c] = make_double2(dv.x * s_delta[c], dv.y * s_delta[c]);
                vD[c] = p.D[s_rows[c]];
            }
            if (p.has_odd) {
                double2 ov = p.oddV[s_rows[sc]];
                vOV[c] = make_double2(ov.x * s_delta[c], ov.y * s_delta[c]);
            }
        }
        __syncthreads();
        // ---- power traces with pairing
        {
            double2 t1 = make_double2(0.0, 0.0);
            for (int r = tid; r < s; r += LG_THREADS) { t1.x += M[r * s + r].x; t1.y += M[r * s + r].y; }
            t1 = block_sum(t1, scratch);
            if (tid == 0) { ptr[0] = make_double2((double)s, 0.0); ptr[1] = t1; }
            double2* Pc = P;
            double2* Pnx = Pn;
            for (int t = 1; 2 * t <= tr_order; ++t) {
                double2 e = make_double2(0.0, 0.0);
                for (int idx = tid; idx < s * s; idx += LG_THREADS) {
                    const int r = idx / s, c = idx % s;
                    cfma2(e, Pc[idx], Pc[c * s + r]);
                }
                e = block_sum(e, scratch);
                if (tid == 0) ptr[2 * t] = e;
                if (2 * t + 1 > tr_order) break;
                for (int idx = tid; idx < s * s; idx += LG_THREADS) {
                    const int r = idx / s, c = idx % s;
                    double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
                    int q = 0;
                    for (; q + 1 < s; q += 2) {
                        cfma2(a0, Pc[r * s + q], M[q * s + c]);
                        cfma2(a1, Pc[r * s + q + 1], M[(q + 1) * s + c]);
                    }
                    if (q < s) cfma2(a0, Pc[r * s + q], M[q * s + c]);
                    Pnx[idx] = make_double2(a0.x + a1.x, a0.y + a1.y);
                }
                __syncthreads();
                double2 o = make_double2(0.0, 0.0);
                for (int idx = tid; idx < s * s; idx += LG_THREADS) {
                    const int r = idx / s, c = idx % s;
                    cfma2(o, Pnx[idx], Pc[c * s + r]);
                }
                o = block_sum(o, scratch);
                if (tid == 0) ptr[2 * t + 1] = o;
                double2* tmp = Pc; Pc = Pnx; Pnx = tmp;
            }
        }
        __syncthreads();
        // ---- series factors fac[i] = i * a_i
        if (!p.has_odd) {
            // a_i = p_i/(2i) + (XD M^(i-1) D)/2   (f / f_loop)
            for (int i = 1; i <= order; ++i) {
                double2 l = make_double2(0.0, 0.0);
                if (loops) {
                    for (int c = tid; c < s; c += LG_THREADS) cfma2(l, vXD[c], vD[c]);
                    l = block_sum(l, scratch);
                    // vD <- M vD
                    for (int r = tid; r < s; r += LG_THREADS) {
                        double2 a = make_double2(0.0, 0.0);
                        for (int q = 0; q < s; ++q) cfma2(a, M[r * s + q], vD[q]);
                        vD2[r] = a;
                    }
                    __syncthreads();
                    for (int r = tid; r < s; r += LG_THREADS) vD[r] = vD2[r];
                    __syncthreads();
                }
                if (tid == 0) fac[i] = make_double2(0.5 * ptr[i].x + 0.5 * i * l.x, 0.5 * ptr[i].y + 0.5 * i * l.y);
            }
        } else {
            // f_loop_odd: a_1 = oddloop; a_{2t} = p_t/(2t) + XD.(M^(t-1) D)/2; a_{2t+1} = oddVX.(M^(t-1) D)
            if (tid == 0) fac[1] = p.oddloop;
            for (int i = 2; i <= order; ++i) {
                if ((i & 1) == 0) {
                    double2 l = make_double2(0.0, 0.0);
                    for (int c = tid; c < s; c += LG_THREADS) cfma2(l, vXD[c], vD[c]);
                    l = block_sum(l, scratch);
                    if (tid == 0) {
                        const int t = i / 2;
                        fac[i] = make_double2(ptr[t].x + 0.5 * i * l.x, ptr[t].y + 0.5 * i * l.y);  // i*(p_t/i + l/2)
                    }
                } else {
                    double2 l = make_double2(0.0, 0.0);
                    for (int c = tid; c < s; c += LG_THREADS) cfma2(l, vOV[c], vD[c]);
                    l = block_sum(l, scratch);
                    if (tid == 0) fac[i] = make_double2(i * l.x, i * l.y);
                    for (int r = tid; r < s; r += LG_THREADS) {
                        double2 a = make_double2(0.0, 0.0);
                        for (int q = 0; q < s; ++q) cfma2(a, M[r * s + q], vD[q]);
                        vD2[r] = a;
                    }
                    __syncthreads();
                    for (int r = tid; r < s; r += LG_THREADS) vD[r] = vD2[r];
                    __syncthreads();
                }
            }
        }
        __syncthreads();
        // ---- c_t = (1/t) sum_{i=1..t} fac[i] c_{t-i}  (warp 0)
        if (tid < 32) {
            if (tid == 0) cser[0] = make_double2(1.0, 0.0);
            __syncwarp();
            for (int t = 1; t <= order; ++t) {
                double2 a = make_double2(0.0, 0.0);
                for (int i = 1 + tid; i <= t; i += 32) cfma2(a, fac[i], cser[t - i]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { a.x += shfl_xor_d(a.x, off); a.y += shfl_xor_d(a.y, off); }
                if (tid == 0) cser[t] = make_double2(a.x / t, a.y / t);
                __syncwarp();
            }
            if (tid == 0) {
                double pre = (((N / 2 - s_esum) & 1) ? -1.0 : 1.0) * s_weight;
                if (p.glynn && !p.has_odd && s_d0zero) pre *= 0.5;
                dd_add(acc.re, pre * cser[order].x);
                dd_add(acc.im, pre * cser[order].y);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        partials[blockIdx.x * 4 + 0] = acc.re.hi; partials[blockIdx.x * 4 + 1] = acc.re.lo;
        partials[blockIdx.x * 4 + 2] = acc.im.hi; partials[blockIdx.x * 4 + 3] = acc.im.lo;
    }
}

static size_t lg_smem_bytes(int n) {
    return sizeof(double2) * ((size_t)3 * n * n + 4 * n + 3 * (LG_MAX_ORDER + 2) + 8);
}

struct DevBufLg {
    void* p = nullptr;
    ~DevBufLg() { if (p) pool_free(p); }
};

}  // namespace wb

using namespace wb;

extern "C" int wb200_lhaf_general_steps(const int32_t* edge_reps, int n_edges, int glynn, int has_odd,
                                        uint64_t* steps) {
    if (!edge_reps || !steps || n_edges < 1) { set_error("lhaf_general_steps: bad arguments"); return WB200_EINVAL; }
    unsigned __int128 s = 1;
    for (int i = 0; i < n_edges; ++i) {
        if (edge_reps[i] < 0) { set_error("negative edge repetition"); return WB200_EINVAL; }
        uint64_t f = (uint64_t)edge_reps[i] + 1;
        if (i == 0 && glynn && !has_odd) f = ((uint64_t)edge_reps[0] + 2) / 2;
        s *= f;
        if (s > (((unsigned __int128)1) << 63)) { set_error("subset index space exceeds 2^63"); return WB200_ENOSUP; }
    }
    *steps = (uint64_t)s;
    return WB200_OK;
}

extern "C" int wb200_lhaf_general_dev(const double* dA, const double* dD, const double* doddV, const double* oddloop,
                                      int n, const int32_t* edge_reps, int glynn, uint64_t j0, uint64_t j1,
                                      double* d_out4, void* stream) {
    if (!dA || !edge_reps || !d_out4) { set_error("lhaf_general: null pointer"); return WB200_EINVAL; }
    if (n < 2 || (n & 1)) { set_error("lhaf_general: n must be even and >= 2 (got %d)", n); return WB200_EINVAL; }
    const int E = n / 2;
    if (E > LG_MAX_EDGES) { set_error("lhaf_general: %d edges exceed the limit of %d", E, LG_MAX_EDGES); return WB200_ENOSUP; }
    const int has_odd = doddV != nullptr;
    if (has_odd && (!oddloop || !dD)) { set_error("lhaf_general: odd vertex needs oddloop and D"); return WB200_EINVAL; }
    uint64_t steps = 0;
    int rc = wb200_lhaf_general_steps(edge_reps, E, glynn, has_odd, &steps);
    if (rc) return rc;
    if (j0 > j1 || j1 > steps) { set_error("lhaf_general: bad subset range"); return WB200_EINVAL; }
    LgParams p;
    memset(&p, 0, sizeof(p));
    int N = has_odd ? 1 : 0;
    for (int i = 0; i < E; ++i) { p.reps[i] = edge_reps[i]; N += 2 * edge_reps[i]; }
    if ((has_odd ? N : N / 2) > LG_MAX_ORDER) { set_error("lhaf_general: photon number %d too large", N); return WB200_ENOSUP; }
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    (void)cudaGetLastError();
    WB_CUDA(cudaGetDevice(&device));
    if (has_odd) p.oddloop = make_double2(oddloop[0], oddloop[1]);
    p.A = (const double2*)dA; p.D = (const double2*)dD; p.oddV = (const double2*)doddV;
    p.n = n; p.E = E; p.glynn = glynn; p.has_odd = has_odd; p.N = N; p.j0 = j0; p.j1 = j1;
    int sms = 0, occ = 1;
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    const size_t shm = lg_smem_bytes(n);
    WB_CUDA(cudaFuncSetAttribute(lhaf_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lhaf_general_kernel, LG_THREADS, shm));
    if (occ < 1) occ = 1;
    uint64_t total = j1 - j0, maxgrid = (uint64_t)sms * occ;
    int grid = (int)(total < maxgrid ? (total ? total : 1) : maxgrid);
    StreamBuf dpart;
    WB_POOL(dpart.alloc(sizeof(double) * 4 * grid, st));
    lhaf_general_kernel<<<grid, LG_THREADS, shm, st>>>(p, (double*)dpart.p);
    final_reduce_kernel<<<1, 32, 0, st>>>((const double*)dpart.p, grid, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

extern "C" int wb200_lhaf_general_host(int device, const double* A, const double* D, const double* oddV,
                                       const double* oddloop, int n, const int32_t* edge_reps, int glynn,
                                       uint64_t j0, uint64_t j1, double out4[4], double* kernel_ms) {
    if (!A || !edge_reps || !out4) { set_error("lhaf_general: null pointer"); return WB200_EINVAL; }
    if (n < 2 || n > 2 * LG_MAX_EDGES) { set_error("lhaf_general: n = %d outside [2, %d]", n, 2 * LG_MAX_EDGES); return n > 2 * LG_MAX_EDGES ? WB200_ENOSUP : WB200_EINVAL; }
    WB_CUDA(cudaSetDevice(device));
    DevBufLg dA, dD, dV, dout;
    WB_POOL(pool_alloc(&dA.p, sizeof(double) * 2 * n * n));
    WB_CUDA(cudaMemcpy(dA.p, A, sizeof(double) * 2 * n * n, cudaMemcpyHostToDevice));
    if (D) {
        WB_POOL(pool_alloc(&dD.p, sizeof(double) * 2 * n));
        WB_CUDA(cudaMemcpy(dD.p, D, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    }
    if (oddV) {
        WB_POOL(pool_alloc(&dV.p, sizeof(double) * 2 * n));
        WB_CUDA(cudaMemcpy(dV.p, oddV, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    }
    WB_POOL(pool_alloc(&dout.p, sizeof(double) * 4));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) {
        WB_CUDA(cudaEventCreate(&e0));
        WB_CUDA(cudaEventCreate(&e1));
        WB_CUDA(cudaEventRecord(e0, 0));
    }
    int rc = wb200_lhaf_general_dev((const double*)dA.p, (const double*)dD.p, (const double*)dV.p, oddloop, n, edge_reps, glynn,
                                    j0, j1, (double*)dout.p, nullptr);
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
