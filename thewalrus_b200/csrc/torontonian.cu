// Torontonian: sum over S subset of [N] of (-1)^(N-|S|) / sqrt(det(I - O_S)).
//
// Replaces rec_torontonian / recursiveTor / quad_cholesky (thewalrus/_torontonian.py:157-247) and the flat
// numba_tor loop (:123-154).  Both reference variants factor every principal submatrix; here the 2^N
// principal minors of B = I - O (modes interleaved as in rec_torontonian :238-241) are enumerated as a
// binary tree of Schur complements: at mode i a node either EXCLUDES the mode (drop its two rows/columns,
// sign *= -1) or INCLUDES it (two scalar pivots d1, d2 of a right-looking Hermitian Cholesky, det *= d1 d2,
// trailing matrix <- Schur complement).  This is the same elimination order as a Cholesky factorisation of
// each B_S, so it is as stable as the reference, but every node only touches the *remaining* modes, whose
// matrices shrink towards the leaves where almost all of the 2^N nodes live.
//
// Work split: the first P = N - DC modes form a "prefix" (the C-ABI range unit).  A CTA owns G consecutive
// prefixes, eliminates their common leading modes once, then for each prefix expands the last DC modes
// breadth-first in shared memory (all threads busy at every level) down to 2-mode (4x4) nodes, which single
// threads finish in registers.
#include "common.cuh"

namespace wb {

constexpr int TOR_THREADS = 256;
constexpr int TOR_DC = 9;        // modes expanded breadth-first inside a CTA
constexpr int TOR_G = 5;         // log2(prefixes per CTA)
constexpr int TOR_MAX_MODES = 32;
constexpr int TOR_BUF = 2304;       // max over BFS levels of nodes * dim^2, DC = 9: 64 nodes of 6 x 6
constexpr int TOR_BUF_LOOP = 3200;  // bordered nodes: 128 nodes of 5 x 5

__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}

// interleave modes and form B = I - O.  Loop torontonian (gamma != nullptr): B is bordered by one extra
// row/column  B[2N][c] = gamma_c, B[c][2N] = conj(gamma_c), B[2N][2N] = 0.  Eliminating a pivot k of the
// bordered Hermitian matrix updates the border row exactly like the forward substitution of the reference
// (solve_triangular, thewalrus/_torontonian.py:250-273: x_i -= L_ik z_k with x = conj(gamma)) and subtracts
// |x_k|^2 / d_k from the corner, so after the kept modes are eliminated  -corner = x^H (I - O_S)^-1 x,
// the exponent of recursiveLTor (:307-309) / numba_ltor (:404), and excluded modes drop out of the border
// together with their rows.  The border index is never a pivot.
__global__ void tor_prep_kernel(const double2* __restrict__ O, const double2* __restrict__ gamma, int N,
                                double2* __restrict__ B) {
    const int n2 = 2 * N, ld = n2 + (gamma ? 1 : 0);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ld * ld; idx += gridDim.x * blockDim.x) {
        const int r = idx / ld, c = idx % ld;
        const int sr = (r >> 1) + (r & 1) * N, sc = (c >> 1) + (c & 1) * N;
        double2 out;
        if (r < n2 && c < n2) {
            const double2 v = O[(size_t)sr * n2 + sc];
            out = make_double2((r == c ? 1.0 : 0.0) - v.x, -v.y);
        } else if (r == n2 && c == n2) {
            out = make_double2(0.0, 0.0);
        } else if (r == n2) {
            out = gamma[sc];
        } else {
            const double2 g = gamma[sr];
            out = make_double2(g.x, -g.y);
        }
        B[idx] = out;
    }
}

// In-place elimination of the leading mode (rows/cols off, off+1) of the dim x dim matrix T (stride ld).
// Returns d1 * d2.  All threads of the CTA participate.
__device__ double eliminate_mode(double2* T, int ld, int off, int dim) {
    const int tid = threadIdx.x;
    // pivot 1
    const double d1 = T[off * ld + off].x;
    const double i1 = 1.0 / d1;
    __syncthreads();
    for (int idx = tid; idx < (dim - off - 1) * (dim - off - 1); idx += TOR_THREADS) {
        const int r = off + 1 + idx / (dim - off - 1), c = off + 1 + idx % (dim - off - 1);
        if (c == off + 1 || r >= off + 2) {  // column off+1 (needed for pivot 2) and the trailing block
            double2 a = T[r * ld + off], b = T[c * ld + off];
            double2 pr = cmulc(a, b);
            T[r * ld + c].x -= pr.x * i1;
            T[r * ld + c].y -= pr.y * i1;
        }
    }
    __syncthreads();
    const double d2 = T[(off + 1) * ld + off + 1].x;
    const double i2 = 1.0 / d2;
    __syncthreads();
    for (int idx = tid; idx < (dim - off - 2) * (dim - off - 2); idx += TOR_THREADS) {
        const int r = off + 2 + idx / (dim - off - 2), c = off + 2 + idx % (dim - off - 2);
        double2 a = T[r * ld + off + 1], b = T[c * ld + off + 1];
        double2 pr = cmulc(a, b);
        T[r * ld + c].x -= pr.x * i2;
        T[r * ld + c].y -= pr.y * i2;
    }
    __syncthreads();
    return d1 * d2;
}

struct TorParams {
    const double2* B;   // interleaved I - O, 2N x 2N
    int N, P, g, DC;    // modes, prefix modes, log2 prefixes per CTA, BFS modes
    uint64_t p0, p1;
};

// finish a 2-mode (4x4 Hermitian) node in registers: 4 subsets
__device__ __forceinline__ double tail2(const double2* T, double det, double sgn) {
    const double t00 = T[0].x, t11 = T[5].x, t22 = T[10].x, t33 = T[15].x;
    const double2 t10 = T[4], t20 = T[8], t30 = T[12], t21 = T[9], t31 = T[13], t32 = T[14];
    double sum = sgn / sqrt(det);                                    // {}: two exclusions, sign unchanged
    const double detB = t22 * t33 - (t32.x * t32.x + t32.y * t32.y);  // {m1}
    sum -= sgn / sqrt(det * detB);
    const double i1 = 1.0 / t00;
    const double d2 = t11 - (t10.x * t10.x + t10.y * t10.y) * i1;
    sum -= sgn / sqrt(det * t00 * d2);                               // {m0}
    // {m0, m1}: Schur complement of mode 0 on the m1 block
    const double i2 = 1.0 / d2;
    double2 u2 = t21, u3 = t31;   // column 1 after pivot 1
    { double2 q = cmulc(t20, t10); u2.x -= q.x * i1; u2.y -= q.y * i1; }
    { double2 q = cmulc(t30, t10); u3.x -= q.x * i1; u3.y -= q.y * i1; }
    const double s22 = t22 - (t20.x * t20.x + t20.y * t20.y) * i1 - (u2.x * u2.x + u2.y * u2.y) * i2;
    const double s33 = t33 - (t30.x * t30.x + t30.y * t30.y) * i1 - (u3.x * u3.x + u3.y * u3.y) * i2;
    double2 s32 = t32;
    { double2 q = cmulc(t30, t20); s32.x -= q.x * i1; s32.y -= q.y * i1; }
    { double2 q = cmulc(u3, u2); s32.x -= q.x * i2; s32.y -= q.y * i2; }
    const double detS = s22 * s33 - (s32.x * s32.x + s32.y * s32.y);
    sum += sgn / sqrt(det * t00 * d2 * detS);
    return sum;
}

// one scalar pivot of a 5x5 bordered node held in registers (lower triangle, index 4 = border)
template <int K>
__device__ __forceinline__ double pivot5(double2 (&a)[5][5]) {
    const double d = a[K][K].x, inv = 1.0 / d;
#pragma unroll
    for (int r = K + 1; r < 5; ++r)
#pragma unroll
        for (int c = K + 1; c <= r; ++c) {
            const double2 q = cmulc(a[r][K], a[c][K]);
            a[r][c].x -= q.x * inv;
            a[r][c].y -= q.y * inv;
        }
    return d;
}

// loop torontonian: finish a 2-mode bordered (5x5) node, 4 subsets; exponent = -corner / 2
__device__ __forceinline__ double tail2_loop(const double2* T, double det, double sgn) {
    double2 a[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = T[r * 5 + c];
    double sum = sgn * exp(-0.5 * a[4][4].x) / sqrt(det);                     // {}
    {                                                                         // {m1}: pivots 2, 3
        const double d1 = a[2][2].x, i1 = 1.0 / d1;
        const double2 e = a[3][2], g2 = a[4][2];
        const double d2 = a[3][3].x - (e.x * e.x + e.y * e.y) * i1;
        double2 u = a[4][3];
        { const double2 q = cmulc(g2, e); u.x -= q.x * i1; u.y -= q.y * i1; }
        const double corner = a[4][4].x - (g2.x * g2.x + g2.y * g2.y) * i1 - (u.x * u.x + u.y * u.y) / d2;
        sum -= sgn * exp(-0.5 * corner) / sqrt(det * d1 * d2);
    }
    const double p0 = pivot5<0>(a);
    const double p1 = pivot5<1>(a);
    const double det01 = det * p0 * p1;
    sum -= sgn * exp(-0.5 * a[4][4].x) / sqrt(det01);                         // {m0}
    const double p2 = pivot5<2>(a);
    const double p3 = pivot5<3>(a);
    sum += sgn * exp(-0.5 * a[4][4].x) / sqrt(det01 * p2 * p3);               // {m0, m1}
    return sum;
}

// AUG = 0: torontonian; AUG = 1: loop torontonian (every matrix carries the border row/column).
template <int AUG>
__global__ void __launch_bounds__(TOR_THREADS) tor_kernel(TorParams p, double* __restrict__ partials) {
    extern __shared__ __align__(16) double smem_tor[];
    const int N = p.N, n2 = 2 * N + AUG, DC = p.DC, g = p.g, P = p.P;
    const int dg = 2 * (DC + g) + AUG;
    double2* T = reinterpret_cast<double2*>(smem_tor);   // n2 x n2
    double2* T0 = T + n2 * n2;                            // dg x dg
    double2* W = T0 + dg * dg;                            // dg x dg
    double2* bufA = W + dg * dg;
    constexpr int bufsz = AUG ? TOR_BUF_LOOP : TOR_BUF;   // max level size for DC = 9 (see host)
    double2* bufB = bufA + bufsz;
    double* detA = reinterpret_cast<double*>(bufB + bufsz);  // 256 each: det, sgn ping-pong; inv1, inv2
    double* sgnA = detA + 256;
    double* detB = sgnA + 256;
    double* sgnB = detB + 256;
    double* inv1 = sgnB + 256;
    double* inv2 = inv1 + 256;
    const int tid = threadIdx.x;

    dd acc = {0.0, 0.0};
    const uint64_t ngroups_first = p.p0 >> g, ngroups_last = (p.p1 + (1ull << g) - 1) >> g;
    for (uint64_t grp = ngroups_first + blockIdx.x; grp < ngroups_last; grp += gridDim.x) {
        // ---- phase A: common leading modes 0 .. P-g-1 (bits of grp, most significant = mode 0)
        __syncthreads();
        for (int idx = tid; idx < n2 * n2; idx += TOR_THREADS) T[idx] = p.B[idx];
        __syncthreads();
        double det0 = 1.0, sgn0 = 1.0;
        const int lead = P - g;
        for (int i = 0; i < lead; ++i) {
            const bool inc = (grp >> (lead - 1 - i)) & 1ull;
            if (inc) det0 *= eliminate_mode(T, n2, 2 * i, n2);
            else sgn0 = -sgn0;
        }
        __syncthreads();
        for (int idx = tid; idx < dg * dg; idx += TOR_THREADS) {
            const int r = idx / dg, c = idx % dg;
            T0[idx] = T[(2 * lead + r) * n2 + 2 * lead + c];
        }
        __syncthreads();
        // ---- phase B: the G prefixes of this group
        for (int sub = 0; sub < (1 << g); ++sub) {
            const uint64_t pfx = (grp << g) + sub;
            if (pfx < p.p0 || pfx >= p.p1) continue;
            __syncthreads();
            for (int idx = tid; idx < dg * dg; idx += TOR_THREADS) W[idx] = T0[idx];
            __syncthreads();
            double det = det0, sgn = sgn0;
            for (int i = 0; i < g; ++i) {
                const bool inc = (sub >> (g - 1 - i)) & 1;
                if (inc) det *= eliminate_mode(W, dg, 2 * i, dg);
                else sgn = -sgn;
            }
            __syncthreads();
            // level 0 node: trailing 2DC x 2DC block
            int dim = 2 * DC + AUG;
            for (int idx = tid; idx < dim * dim; idx += TOR_THREADS) {
                const int r = idx / dim, c = idx % dim;
                bufA[idx] = W[(2 * g + r) * dg + 2 * g + c];
            }
            if (tid == 0) { detA[0] = det; sgnA[0] = sgn; }
            __syncthreads();
            double2* cur = bufA; double2* nxt = bufB;
            double* dcur = detA; double* scur = sgnA; double* dnxt = detB; double* snxt = sgnB;
            int nodes = 1;
            while (dim > 4 + AUG) {
                const int cd = dim - 2;
                // per-node pivots
                for (int nd = tid; nd < nodes; nd += TOR_THREADS) {
                    const double2* Tn = cur + nd * dim * dim;
                    const double d1 = Tn[0].x, i1 = 1.0 / d1;
                    const double2 e = Tn[dim];  // T[1][0]
                    const double d2 = Tn[dim + 1].x - (e.x * e.x + e.y * e.y) * i1;
                    inv1[nd] = i1; inv2[nd] = 1.0 / d2;
                    dnxt[2 * nd] = dcur[nd]; snxt[2 * nd] = -scur[nd];            // exclude
                    dnxt[2 * nd + 1] = dcur[nd] * d1 * d2; snxt[2 * nd + 1] = scur[nd];  // include
                }
                __syncthreads();
                const int per = cd * cd;
                for (int idx = tid; idx < nodes * 2 * per; idx += TOR_THREADS) {
                    const int child = idx / per, el = idx % per;
                    const int nd = child >> 1, r = el / cd + 2, c = el % cd + 2;
                    const double2* Tn = cur + nd * dim * dim;
                    double2 v = Tn[r * dim + c];
                    if (child & 1) {
                        const double i1 = inv1[nd], i2 = inv2[nd];
                        const double2 a = Tn[r * dim], b = Tn[c * dim], e = Tn[dim];
                        double2 ur = Tn[r * dim + 1], uc = Tn[c * dim + 1];
                        { double2 q = cmulc(a, e); ur.x -= q.x * i1; ur.y -= q.y * i1; }
                        { double2 q = cmulc(b, e); uc.x -= q.x * i1; uc.y -= q.y * i1; }
                        { double2 q = cmulc(a, b); v.x -= q.x * i1; v.y -= q.y * i1; }
                        { double2 q = cmulc(ur, uc); v.x -= q.x * i2; v.y -= q.y * i2; }
                    }
                    nxt[child * per + el] = v;
                }
                __syncthreads();
                { double2* t = cur; cur = nxt; nxt = t; }
                { double* t = dcur; dcur = dnxt; dnxt = t; t = scur; scur = snxt; snxt = t; }
                nodes *= 2; dim = cd;
            }
            // ---- 2-mode nodes finished by single threads
            for (int nd = tid; nd < nodes; nd += TOR_THREADS) {
                if (AUG) dd_add(acc, tail2_loop(cur + nd * 25, dcur[nd], scur[nd]));
                else dd_add(acc, tail2(cur + nd * 16, dcur[nd], scur[nd]));
            }
        }
    }
    __shared__ double red[(TOR_THREADS / 32) * 4];
    cdd a2;
    a2.re = acc;
    a2.im = {0.0, 0.0};
    block_reduce_store(a2, red, partials);
}

static void tor_shape(int N, int* P, int* g, int* DC) {
    *DC = N < TOR_DC ? N : TOR_DC;
    *P = N - *DC;
    *g = *P < TOR_G ? *P : TOR_G;
}

struct DevBufT {
    void* p = nullptr;
    ~DevBufT() { if (p) pool_free(p); }
};

}  // namespace wb

using namespace wb;

extern "C" int wb200_tor_num_prefixes(int n_modes, uint64_t* count) {
    if (!count || n_modes < 2 || n_modes > TOR_MAX_MODES) {
        set_error("tor: number of modes %d outside [2, %d]", n_modes, TOR_MAX_MODES);
        return (n_modes > TOR_MAX_MODES) ? WB200_ENOSUP : WB200_EINVAL;
    }
    int P, g, DC;
    tor_shape(n_modes, &P, &g, &DC);
    *count = 1ull << P;
    return WB200_OK;
}

// workspace: interleaved (bordered) I - O, (2N + 1) x (2N + 1) complex, followed by 4 doubles per CTA
static size_t tor_ws_partials_offset(int N) {
    return (sizeof(double2) * (size_t)(2 * N + 1) * (2 * N + 1) + 255) & ~(size_t)255;
}
constexpr int TOR_MAX_GRID = 1024;

extern "C" size_t wb200_tor_workspace_bytes(int n_modes) {
    if (n_modes < 2 || n_modes > TOR_MAX_MODES) return 0;
    return tor_ws_partials_offset(n_modes) + sizeof(double) * 4 * TOR_MAX_GRID;
}

// shared launcher: dGamma == nullptr -> torontonian, else loop torontonian
static int tor_launch(const double* dO, const double* dGamma, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                      void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!dO || !d_out4 || !d_workspace) { set_error("tor: null pointer"); return WB200_EINVAL; }
    uint64_t total = 0;
    int rc = wb200_tor_num_prefixes(n_modes, &total);
    if (rc) return rc;
    if (p0 > p1 || p1 > total) { set_error("tor: bad prefix range"); return WB200_EINVAL; }
    if (workspace_bytes < wb200_tor_workspace_bytes(n_modes)) { set_error("tor: workspace too small"); return WB200_EINVAL; }
    const int N = n_modes, aug = dGamma ? 1 : 0, n2 = 2 * N + aug;
    cudaStream_t st = (cudaStream_t)stream;
    TorParams p;
    tor_shape(N, &p.P, &p.g, &p.DC);
    p.N = N; p.p0 = p0; p.p1 = p1;
    double2* dB = reinterpret_cast<double2*>(d_workspace);
    double* dpart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_workspace) + tor_ws_partials_offset(N));
    p.B = dB;
    int dev = 0, sms = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (device_sm_count(dev, &sms)) return WB200_ECUDA;
    const int dg = 2 * (p.DC + p.g) + aug;
    const size_t shm = sizeof(double2) * ((size_t)n2 * n2 + 2 * (size_t)dg * dg + 2 * (aug ? TOR_BUF_LOOP : TOR_BUF)) +
                       sizeof(double) * 6 * 256;
    {
        const int max_n2 = 2 * TOR_MAX_MODES + 1, max_dg = 2 * (TOR_DC + TOR_G) + 1;
        const int max_shm = (int)(sizeof(double2) * ((size_t)max_n2 * max_n2 + 2 * (size_t)max_dg * max_dg + 2 * TOR_BUF_LOOP) +
                                  sizeof(double) * 6 * 256);
        if (aug) WB_CUDA(cudaFuncSetAttribute(tor_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_shm));
        else WB_CUDA(cudaFuncSetAttribute(tor_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_shm));
    }
    const uint64_t groups = ((p1 + (1ull << p.g) - 1) >> p.g) - (p0 >> p.g);
    int grid = (int)(groups < (uint64_t)sms ? (groups ? groups : 1) : (uint64_t)sms);
    if (grid > TOR_MAX_GRID) grid = TOR_MAX_GRID;
    tor_prep_kernel<<<8, 256, 0, st>>>(reinterpret_cast<const double2*>(dO), reinterpret_cast<const double2*>(dGamma), N, dB);
    if (aug) tor_kernel<1><<<grid, TOR_THREADS, shm, st>>>(p, dpart);
    else tor_kernel<0><<<grid, TOR_THREADS, shm, st>>>(p, dpart);
    final_reduce_kernel<<<1, 32, 0, st>>>(dpart, grid, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

extern "C" int wb200_tor_dev(const double* dO, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
    return tor_launch(dO, nullptr, n_modes, p0, p1, d_out4, d_workspace, workspace_bytes, stream);
}

extern "C" int wb200_ltor_dev(const double* dO, const double* dGamma, int n_modes, uint64_t p0, uint64_t p1,
                              double* d_out4, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!dGamma) { set_error("ltor: null gamma"); return WB200_EINVAL; }
    return tor_launch(dO, dGamma, n_modes, p0, p1, d_out4, d_workspace, workspace_bytes, stream);
}

static int tor_host_impl(int device, const double* O, const double* gamma, int n_modes, uint64_t p0, uint64_t p1,
                         double out2[2], double* kernel_ms) {
    if (!O || !out2) { set_error("tor: null pointer"); return WB200_EINVAL; }
    uint64_t total = 0;
    int rc = wb200_tor_num_prefixes(n_modes, &total);
    if (rc) return rc;
    const int n2 = 2 * n_modes;
    WB_CUDA(cudaSetDevice(device));
    DevBufT dO, dG, dws, dout;
    const size_t wsb = wb200_tor_workspace_bytes(n_modes);
    WB_POOL(pool_alloc(&dO.p, sizeof(double) * 2 * n2 * n2));
    WB_POOL(pool_alloc(&dws.p, wsb));
    WB_POOL(pool_alloc(&dout.p, sizeof(double) * 4));
    WB_CUDA(cudaMemcpy(dO.p, O, sizeof(double) * 2 * n2 * n2, cudaMemcpyHostToDevice));
    if (gamma) {
        WB_POOL(pool_alloc(&dG.p, sizeof(double) * 2 * n2));
        WB_CUDA(cudaMemcpy(dG.p, gamma, sizeof(double) * 2 * n2, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    WB_CUDA(cudaEventCreate(&e0));
    WB_CUDA(cudaEventCreate(&e1));
    WB_CUDA(cudaEventRecord(e0, 0));
    rc = tor_launch((const double*)dO.p, (const double*)dG.p, n_modes, p0, p1, (double*)dout.p, dws.p, wsb, nullptr);
    if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
    WB_CUDA(cudaEventRecord(e1, 0));
    WB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    WB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    double o4[4];
    WB_CUDA(cudaMemcpy(o4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    out2[0] = o4[0]; out2[1] = o4[1];
    return WB200_OK;
}

extern "C" int wb200_tor_host(int device, const double* O, int n_modes, uint64_t p0, uint64_t p1, double out2[2],
                              double* kernel_ms) {
    return tor_host_impl(device, O, nullptr, n_modes, p0, p1, out2, kernel_ms);
}

extern "C" int wb200_ltor_host(int device, const double* O, const double* gamma, int n_modes, uint64_t p0, uint64_t p1,
                               double out2[2], double* kernel_ms) {
    if (!gamma) { set_error("ltor: null gamma"); return WB200_EINVAL; }
    return tor_host_impl(device, O, gamma, n_modes, p0, p1, out2, kernel_ms);
}
