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
// Work split: the first P = N - DC modes form a "prefix" (the C-ABI range unit).  A CTA owns 2^g consecutive
// prefixes: it eliminates their common leading modes once (in place), walks the g group modes depth-first (one
// elimination per prefix), and expands the last DC modes of every prefix breadth-first in shared memory down to
// 2-mode (4x4, loop variant 5x5) nodes, which single threads finish in registers.  Nodes are (offset, stride)
// views of stored lower triangles, so excluding a mode copies nothing; see the comment above tor_kernel.
#include <stdlib.h>
#include "common.cuh"

namespace wb {

#ifndef WB_TOR_THREADS      // build-time overrides for the shape experiment (tools/gpu_tor_shape.py)
#define WB_TOR_THREADS 256  // measured: 256 threads 1.23 ms, 512 threads 1.36 ms, 128 threads 1.42 ms at 2N = 48
#endif
#ifndef WB_TOR_DC
#define WB_TOR_DC 9
#endif
#ifndef WB_TOR_G
#define WB_TOR_G 5
#endif
#ifndef WB_TOR_SMEM_KB
#define WB_TOR_SMEM_KB 225
#endif
constexpr int TOR_THREADS = WB_TOR_THREADS;  // CTA size of the torontonian (profiles/r01_tor_shape.txt)
constexpr int TOR_THREADS_LOOP = 256;  // loop torontonian: the 5 x 5 register tail needs > 128 registers
constexpr int TOR_DC = WB_TOR_DC;      // modes expanded breadth-first inside a CTA
constexpr int TOR_G = WB_TOR_G;        // log2(prefixes per CTA)
constexpr int TOR_MAX_MODES = 32;

__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}

// Scalar type of the tree: double2 (Hermitian I - O) or double (REAL SYMMETRIC O, e.g. real squeezing through a real
// interferometer: a quarter of the multiplications per Schur-complement entry and half the shared memory per node).
template <typename T> struct TorS;
template <> struct TorS<double2> {
    static __device__ __forceinline__ double re(double2 a) { return a.x; }
    static __device__ __forceinline__ double absq(double2 a) { return a.x * a.x + a.y * a.y; }
    static __device__ __forceinline__ double2 mulc(double2 a, double2 b) { return cmulc(a, b); }
    static __device__ __forceinline__ void subs(double2& v, double2 q, double s) { v.x -= q.x * s; v.y -= q.y * s; }
};
template <> struct TorS<double> {
    static __device__ __forceinline__ double re(double a) { return a; }
    static __device__ __forceinline__ double absq(double a) { return a * a; }
    static __device__ __forceinline__ double mulc(double a, double b) { return a * b; }
    static __device__ __forceinline__ void subs(double& v, double q, double s) { v = fma(-q, s, v); }
};

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

// real symmetric O (torontonian only): the same interleaving, B = I - O in doubles
__global__ void tor_prep_real_kernel(const double* __restrict__ O, int N, double* __restrict__ B) {
    const int n2 = 2 * N;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n2 * n2; idx += gridDim.x * blockDim.x) {
        const int r = idx / n2, c = idx % n2;
        const int sr = (r >> 1) + (r & 1) * N, sc = (c >> 1) + (c & 1) * N;
        B[idx] = (r == c ? 1.0 : 0.0) - O[(size_t)sr * n2 + sc];
    }
}

// Lower-triangle enumeration: el -> (r, c), r >= c, el = r (r + 1) / 2 + c.
__device__ __forceinline__ void tri_decode(int el, int& r, int& c) {
    int rr = (int)((__fsqrt_rn((float)(8 * el + 1)) - 1.0f) * 0.5f);   // exact up to one unit for el < 2^20
    rr -= (rr * (rr + 1) / 2 > el) ? 1 : 0;
    rr += ((rr + 1) * (rr + 2) / 2 <= el) ? 1 : 0;
    r = rr;
    c = el - rr * (rr + 1) / 2;
}

// 1 / d1 and 1 / d2 of the two pivots of a leading mode (d2 = t11 - |e|^2 / d1) from ONE division:
// with w = d1 t11 - |e|^2 = d1 d2 and q = 1 / (d1 w):  1 / d1 = q w,  1 / d2 = d1 / w = q d1^2.
template <typename T>
__device__ __forceinline__ void pivot_inverses(double d1, double t11, T e, double& i1, double& i2, double& d1d2) {
    const double w = fma(d1, t11, -TorS<T>::absq(e));
    const double q = 1.0 / (d1 * w);
    i1 = q * w;
    i2 = q * d1 * d1;
    d1d2 = w;
}

// One entry of the Schur complement of the leading mode (rows/cols 0, 1) of the Hermitian matrix P (stride ld):
// both scalar pivots fused.  Reads only the lower triangle of P (columns 0, 1 and the entry itself), so it can
// run in place.  r, c >= 2 index P.
template <typename T>
__device__ __forceinline__ T schur_entry(const T* P, int ld, int r, int c, T e, double i1, double i2) {
    using X = TorS<T>;
    T v = P[r * ld + c];
    const T a = P[r * ld], b = P[c * ld];
    T ur = P[r * ld + 1], uc = P[c * ld + 1];
    X::subs(ur, X::mulc(a, e), i1);
    X::subs(uc, X::mulc(b, e), i1);
    X::subs(v, X::mulc(a, b), i1);
    X::subs(v, X::mulc(ur, uc), i2);
    return v;
}

// Eliminate the leading mode of P (dim x dim, stride ld) into Q (stride ldq; Q may be P + 2 (ld + 1) with ldq = ld:
// in place).  Lower triangle only.  All threads of the CTA take part; returns d1 * d2.
template <int THREADS, typename T>
__device__ double eliminate_into(const T* P, int ld, int dim, T* Q, int ldq) {
    __syncthreads();
    const T e = P[ld];
    double i1, i2, d1d2;
    pivot_inverses(TorS<T>::re(P[0]), TorS<T>::re(P[ld + 1]), e, i1, i2, d1d2);
    const int cd = dim - 2, tri = cd * (cd + 1) / 2;
    for (int el = threadIdx.x; el < tri; el += THREADS) {
        int r, c;
        tri_decode(el, r, c);
        Q[r * ldq + c] = schur_entry(P, ld, r + 2, c + 2, e, i1, i2);
    }
    __syncthreads();
    return d1d2;
}

struct TorParams {
    const void* B;      // interleaved (bordered) I - O: double2, or double for the real-symmetric torontonian
    int N, P, g, DC;    // modes, prefix modes, log2 prefixes per CTA, BFS modes
    int off_depth, off_pool, off_desc;   // shared-memory layout, in double2 units from the start
    uint64_t p0, p1;
};

// finish a 2-mode (4x4 Hermitian, stride ld, lower triangle) node in registers: 4 subsets
template <typename TT>
__device__ __forceinline__ double tail2(const TT* T, int ld, double det, double sgn) {
    using X = TorS<TT>;
    const double t00 = X::re(T[0]), t11 = X::re(T[ld + 1]), t22 = X::re(T[2 * ld + 2]), t33 = X::re(T[3 * ld + 3]);
    const TT t10 = T[ld], t20 = T[2 * ld], t30 = T[3 * ld], t21 = T[2 * ld + 1], t31 = T[3 * ld + 1],
             t32 = T[3 * ld + 2];
    double sum = rsqrt(det);                                         // {}: two exclusions, sign unchanged
    const double detB = t22 * t33 - X::absq(t32);                    // {m1}
    sum -= rsqrt(det * detB);
    const double i1 = 1.0 / t00;
    const double d2 = t11 - X::absq(t10) * i1;
    sum -= rsqrt(det * t00 * d2);                                    // {m0}
    // {m0, m1}: Schur complement of mode 0 on the m1 block
    const double i2 = 1.0 / d2;
    TT u2 = t21, u3 = t31;   // column 1 after pivot 1
    X::subs(u2, X::mulc(t20, t10), i1);
    X::subs(u3, X::mulc(t30, t10), i1);
    const double s22 = t22 - X::absq(t20) * i1 - X::absq(u2) * i2;
    const double s33 = t33 - X::absq(t30) * i1 - X::absq(u3) * i2;
    TT s32 = t32;
    X::subs(s32, X::mulc(t30, t20), i1);
    X::subs(s32, X::mulc(u3, u2), i2);
    const double detS = s22 * s33 - X::absq(s32);
    sum += rsqrt(det * t00 * d2 * detS);
    return sgn * sum;
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

// loop torontonian: finish a 2-mode bordered (5x5, stride ld) node, 4 subsets; exponent = -corner / 2
__device__ __forceinline__ double tail2_loop_body(double2 (&a)[5][5], double det, double sgn) {
    double sum = exp(-0.5 * a[4][4].x) * rsqrt(det);                          // {}
    {                                                                         // {m1}: pivots 2, 3
        const double d1 = a[2][2].x, i1 = 1.0 / d1;
        const double2 e = a[3][2], g2 = a[4][2];
        const double d2 = a[3][3].x - (e.x * e.x + e.y * e.y) * i1;
        double2 u = a[4][3];
        { const double2 q = cmulc(g2, e); u.x -= q.x * i1; u.y -= q.y * i1; }
        const double corner = a[4][4].x - (g2.x * g2.x + g2.y * g2.y) * i1 - (u.x * u.x + u.y * u.y) / d2;
        sum -= exp(-0.5 * corner) * rsqrt(det * d1 * d2);
    }
    const double p0 = pivot5<0>(a);
    const double p1 = pivot5<1>(a);
    const double det01 = det * p0 * p1;
    sum -= exp(-0.5 * a[4][4].x) * rsqrt(det01);                              // {m0}
    const double p2 = pivot5<2>(a);
    const double p3 = pivot5<3>(a);
    sum += exp(-0.5 * a[4][4].x) * rsqrt(det01 * p2 * p3);                    // {m0, m1}
    return sgn * sum;
}
__device__ __forceinline__ double tail2_loop(const double2* T, int ld, double det, double sgn) {       // square storage (v4)
    double2 a[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = T[r * ld + c];
    return tail2_loop_body(a, det, sgn);
}
__device__ __forceinline__ double tail2_loop_p(const double2* T, int k, double det, double sgn) {      // packed view (T, k)
    double2 a[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const int rb = ((r + k) * (r + k + 1) >> 1) + k;
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = T[rb + c];
    }
    return tail2_loop_body(a, det, sgn);
}

// ---- packed lower triangles (tor_kernel) -------------------------------------------------------------------------------
// A node of the subset tree is a VIEW (ptr, k) of a stored Hermitian matrix: the packed lower triangle at S + ptr
// (row r starts at r (r + 1) / 2) with the leading k rows and columns dropped.  Element (r, c), r >= c, of the view:
__device__ __forceinline__ int tv_row(int k, int r) { const int rr = r + k; return ((rr * (rr + 1)) >> 1) + k; }
__host__ __device__ __forceinline__ int tor_tri(int dim) { return dim * (dim + 1) / 2; }
// Round 1 stored full squares with an odd stride (tor_ld): twice the shared memory, ONE CTA per SM at 2N = 48.  Packed,
// two CTAs of the complex kernel (four of the real one) share an SM and cover each other's barriers.

// One entry of the Schur complement of the leading mode (rows/cols 0, 1) of the view (P, k): both scalar pivots fused.
// Reads only columns 0, 1 and the entry itself, so it can run in place.  r, c >= 2 index the view.
template <typename T>
__device__ __forceinline__ T schur_entry_p(const T* P, int k, int r, int c, T e, double i1, double i2) {
    using X = TorS<T>;
    const int rb = tv_row(k, r), cb = tv_row(k, c);
    T v = P[rb + c];
    const T a = P[rb], b = P[cb];
    T ur = P[rb + 1], uc = P[cb + 1];
    X::subs(ur, X::mulc(a, e), i1);
    X::subs(uc, X::mulc(b, e), i1);
    X::subs(v, X::mulc(a, b), i1);
    X::subs(v, X::mulc(ur, uc), i2);
    return v;
}

// Eliminate the leading mode of the view (P, k) of dimension dim into the view (Q, kq) (Q == P, kq == k + 2: in place).
// All threads of the CTA take part; returns d1 * d2.
template <int THREADS, typename T>
__device__ double eliminate_into_p(const T* P, int k, int dim, T* Q, int kq) {
    __syncthreads();
    const int r1 = tv_row(k, 1);
    const T e = P[r1];
    double i1, i2, d1d2;
    pivot_inverses(TorS<T>::re(P[tv_row(k, 0)]), TorS<T>::re(P[r1 + 1]), e, i1, i2, d1d2);
    const int cd = dim - 2, tri = cd * (cd + 1) / 2;
    for (int el = threadIdx.x; el < tri; el += THREADS) {
        int r, c;
        tri_decode(el, r, c);
        Q[tv_row(kq, r) + c] = schur_entry_p(P, k, r + 2, c + 2, e, i1, i2);
    }
    __syncthreads();
    return d1d2;
}

// finish a 2-mode (4x4 Hermitian) node (T, k) in registers: 4 subsets
template <typename TT>
__device__ __forceinline__ double tail2_p(const TT* T, int k, double det, double sgn) {
    using X = TorS<TT>;
    const int r0 = tv_row(k, 0), r1 = tv_row(k, 1), r2 = tv_row(k, 2), r3 = tv_row(k, 3);
    const double t00 = X::re(T[r0]), t11 = X::re(T[r1 + 1]), t22 = X::re(T[r2 + 2]), t33 = X::re(T[r3 + 3]);
    const TT t10 = T[r1], t20 = T[r2], t30 = T[r3], t21 = T[r2 + 1], t31 = T[r3 + 1], t32 = T[r3 + 2];
    double sum = rsqrt(det);                                         // {}: two exclusions, sign unchanged
    const double detB = t22 * t33 - X::absq(t32);                    // {m1}
    sum -= rsqrt(det * detB);
    const double i1 = 1.0 / t00;
    const double d2 = t11 - X::absq(t10) * i1;
    sum -= rsqrt(det * t00 * d2);                                    // {m0}
    // {m0, m1}: Schur complement of mode 0 on the m1 block
    const double i2 = 1.0 / d2;
    TT u2 = t21, u3 = t31;   // column 1 after pivot 1
    X::subs(u2, X::mulc(t20, t10), i1);
    X::subs(u3, X::mulc(t30, t10), i1);
    const double s22 = t22 - X::absq(t20) * i1 - X::absq(u2) * i2;
    const double s33 = t33 - X::absq(t30) * i1 - X::absq(u3) * i2;
    TT s32 = t32;
    X::subs(s32, X::mulc(t30, t20), i1);
    X::subs(s32, X::mulc(u3, u2), i2);
    const double detS = s22 * s33 - X::absq(s32);
    sum += rsqrt(det * t00 * d2 * detS);
    return sgn * sum;
}

// Shared-memory plan of tor_kernel (units of the scalar type): T (packed n2) | depth buffers of the prefix DFS | include-
// children pool of the breadth-first expansion | node descriptors.  Returns the total in bytes.
__host__ __device__ inline size_t tor_smem_plan_p(int N, int aug, int g, int DC, int* off_depth, int* off_pool, int* off_desc,
                                                  size_t elem) {
    const int n2 = 2 * N + aug, dg = 2 * (DC + g) + aug;
    int off = tor_tri(n2);
    *off_depth = off;
    for (int k = 1; k <= g; ++k) off += tor_tri(dg - 2 * k);
    *off_pool = off;
    for (int l = 0, dim = 2 * DC + aug; dim > 4 + aug; ++l, dim -= 2) off += (1 << l) * tor_tri(dim - 2);
    off = (off + 1) & ~1;                                            // descriptors (int / double) start 16-byte aligned
    *off_desc = off;
    const size_t desc = (size_t)128 * (2 * 2 * sizeof(int) + 2 * 2 * sizeof(double)) + 2 * 64 * sizeof(double);
    return (size_t)off * elem + desc;
}

// Row stride of a stored dim x dim matrix: odd, so that the column reads P[r * ld] / P[r * ld + 1] of a quarter
// warp (8 lanes x 16 bytes) fall in 8 different 16-byte bank groups instead of one.
__host__ __device__ __forceinline__ int tor_ld(int dim) { return dim | 1; }

constexpr int TOR_MAXG = 5;
constexpr int TOR_MAXNODES = 128;   // 2-mode tail nodes of a 9-mode breadth-first expansion

// Shared-memory plan (units of the scalar type, double2 or double): T (n2^2) | depth buffers of the prefix DFS | include-children pool of the
// breadth-first expansion | node descriptors.  Returns the total in bytes.
__host__ __device__ inline size_t tor_smem_plan(int N, int aug, int g, int DC, int* off_depth, int* off_pool, int* off_desc,
                                                size_t elem = sizeof(double2)) {
    const int n2 = 2 * N + aug, dg = 2 * (DC + g) + aug;
    int off = n2 * tor_ld(n2);
    *off_depth = off;
    for (int k = 1; k <= g; ++k) off += (dg - 2 * k) * tor_ld(dg - 2 * k);
    *off_pool = off;
    for (int l = 0, dim = 2 * DC + aug; dim > 4 + aug; ++l, dim -= 2) off += (1 << l) * (dim - 2) * tor_ld(dim - 2);
    *off_desc = off;
    // descriptors: 2 x (ptr, ld) int + 2 x (det, sgn) double per node, ping-pong, + inv1/inv2 per parent
    const size_t desc = (size_t)TOR_MAXNODES * (2 * 2 * sizeof(int) + 2 * 2 * sizeof(double)) + 2 * 64 * sizeof(double);
    return (((size_t)off * elem + 15) & ~(size_t)15) + desc;
}

// AUG = 0: torontonian; AUG = 1: loop torontonian (every matrix carries the border row/column).
//
// v2 of the tree walk.  Every node of the subset tree is a VIEW (offset, stride) of some stored Hermitian matrix
// of which only the lower triangle is valid:
//   * excluding the leading mode is free: the child is the parent's view advanced by two rows and columns;
//   * including it writes the Schur complement (both pivots fused, lower triangle only) into fresh storage.
// The first P - g modes (bits of the group id) are walked in place on T.  The next g modes select one of 2^g
// prefixes; they are visited in index order as a depth-first walk with one buffer per depth, so stepping to the
// next prefix costs ONE elimination (the level whose bit turns 0 -> 1; the levels below restart as exclusions).
// The last DC modes are expanded breadth-first: level l has 2^l nodes, each served by 256 / 2^l threads, the
// included children go to a pool that stays alive until the 2-mode tails have been summed in registers.
template <int AUG, int THREADS, typename T = double2>
__global__ void __launch_bounds__(THREADS) tor_kernel(TorParams p, double* __restrict__ partials) {
    static_assert(AUG == 0 || sizeof(T) == sizeof(double2), "the loop torontonian has a complex border: complex only");
    constexpr int LOG_THREADS = THREADS == 512 ? 9 : (THREADS == 256 ? 8 : 7);
    static_assert(THREADS == 128 || THREADS == 256 || THREADS == 512, "power of two, at least one thread per parent node");
    extern __shared__ __align__(16) double smem_tor[];
    const int N = p.N, n2 = 2 * N + AUG, DC = p.DC, g = p.g, P = p.P;
    const int dg = 2 * (DC + g) + AUG;
    T* S = reinterpret_cast<T*>(smem_tor);
    const T* Bg = static_cast<const T*>(p.B);
    int* ptrA = reinterpret_cast<int*>(S + p.off_desc);
    int* ldA = ptrA + TOR_MAXNODES;          // (the k of a view: leading rows / columns dropped)
    int* ptrB = ldA + TOR_MAXNODES;
    int* ldB = ptrB + TOR_MAXNODES;
    double* detA = reinterpret_cast<double*>(ldB + TOR_MAXNODES);
    double* sgnA = detA + TOR_MAXNODES;
    double* detB = sgnA + TOR_MAXNODES;
    double* sgnB = detB + TOR_MAXNODES;
    const int tid = threadIdx.x;
    __shared__ uchar2 tri_rc[160];   // el -> (r, c) of the lower triangle, up to 17 x 17 (the largest breadth-first child)
    for (int el = tid; el < 160; el += THREADS) {
        int r, c;
        tri_decode(el, r, c);
        tri_rc[el] = make_uchar2((unsigned char)r, (unsigned char)c);
    }

    dd acc = {0.0, 0.0};
    const uint64_t ngroups_first = p.p0 >> g, ngroups_last = (p.p1 + (1ull << g) - 1) >> g;
    for (uint64_t grp = ngroups_first + blockIdx.x; grp < ngroups_last; grp += gridDim.x) {
        // ---- phase A: common leading modes 0 .. P-g-1 (bits of grp, most significant = mode 0), in place on T
        __syncthreads();
        for (int el = tid; el < tor_tri(n2); el += THREADS) {      // lower triangle of the (bordered) I - O, packed
            int r, c;
            tri_decode(el, r, c);
            S[el] = Bg[r * n2 + c];
        }
        double det0 = 1.0, sgn0 = 1.0;
        const int lead = P - g;
        for (int i = 0; i < lead; ++i) {
            const bool inc = (grp >> (lead - 1 - i)) & 1ull;
            if (inc) det0 *= eliminate_into_p<THREADS>(S, 2 * i, n2 - 2 * i, S, 2 * i + 2);     // in place
            else sgn0 = -sgn0;
        }
        // ---- phase B: the 2^g prefixes of this group, depth-first
        int nptr[TOR_MAXG + 1], nld[TOR_MAXG + 1];
        double ndet[TOR_MAXG + 1], nsgn[TOR_MAXG + 1];
        nptr[0] = 0; nld[0] = 2 * lead; ndet[0] = det0; nsgn[0] = sgn0;       // (nptr, nld) = the view (ptr, k)
#pragma unroll
        for (int k = 1; k <= TOR_MAXG; ++k) { nptr[k] = 0; nld[k] = 0; ndet[k] = 1.0; nsgn[k] = 1.0; }
        int cur = -1;
        for (int sub = 0; sub < (1 << g); ++sub) {
            const uint64_t pfx = (grp << g) + sub;
            if (pfx < p.p0 || pfx >= p.p1) continue;
            // levels before `first` are shared with the previously materialised prefix
            const int first = cur < 0 ? 0 : g - (32 - __clz(sub ^ cur));
            cur = sub;
            int dbuf = p.off_depth;
#pragma unroll
            for (int lvl = 0; lvl < TOR_MAXG; ++lvl) {
                if (lvl < g) {
                    const int dim = dg - 2 * lvl, cdim = dim - 2;
                    if (lvl >= first) {
                        const bool inc = (sub >> (g - 1 - lvl)) & 1;
                        if (inc) {
                            ndet[lvl + 1] = ndet[lvl] * eliminate_into_p<THREADS>(S + nptr[lvl], nld[lvl], dim, S + dbuf, 0);
                            nsgn[lvl + 1] = nsgn[lvl];
                            nptr[lvl + 1] = dbuf; nld[lvl + 1] = 0;
                        } else {
                            ndet[lvl + 1] = ndet[lvl]; nsgn[lvl + 1] = -nsgn[lvl];
                            nptr[lvl + 1] = nptr[lvl]; nld[lvl + 1] = nld[lvl] + 2;
                        }
                    }
                    dbuf += tor_tri(cdim);
                }
            }
            // ---- breadth-first expansion of the last DC modes
            int rptr = nptr[0], rld = nld[0];
            double rdet = ndet[0], rsgn = nsgn[0];
#pragma unroll
            for (int k = 1; k <= TOR_MAXG; ++k)
                if (k == g) { rptr = nptr[k]; rld = nld[k]; rdet = ndet[k]; rsgn = nsgn[k]; }
            __syncthreads();   // previous prefix's tails are done with the descriptors and the pool
            if (tid == 0) { ptrA[0] = rptr; ldA[0] = rld; detA[0] = rdet; sgnA[0] = rsgn; }
            int* pc = ptrA; int* lc = ldA; double* dc = detA; double* sc = sgnA;
            int* pn = ptrB; int* ln = ldB; double* dn = detB; double* sn = sgnB;
            int nodes = 1, shift = LOG_THREADS, pool = p.off_pool;
            __syncthreads();
            for (int dim = 2 * DC + AUG; dim > 4 + AUG; dim -= 2) {
                const int cd = dim - 2, tri = cd * (cd + 1) / 2, csz = tri;
                const int nd = tid >> shift, lane = tid & ((1 << shift) - 1), step = 1 << shift;
                const int myptr = pc[nd], myld = lc[nd];
                const T* Pn = S + myptr;
                T* Qn = S + pool + nd * csz;
                // every thread of the node derives the two pivots itself: no separate pivot pass, one barrier per level
                const int r1 = tv_row(myld, 1);
                const T e = Pn[r1];
                double i1, i2, d1d2;
                pivot_inverses(TorS<T>::re(Pn[tv_row(myld, 0)]), TorS<T>::re(Pn[r1 + 1]), e, i1, i2, d1d2);
                if (lane == 0) {
                    const double dt = dc[nd], sg = sc[nd];
                    pn[2 * nd] = myptr; ln[2 * nd] = myld + 2; dn[2 * nd] = dt; sn[2 * nd] = -sg;   // exclude: the same storage, two more rows dropped
                    pn[2 * nd + 1] = pool + nd * csz; ln[2 * nd + 1] = 0; dn[2 * nd + 1] = dt * d1d2; sn[2 * nd + 1] = sg;
                }
                for (int el = lane; el < tri; el += step) {
                    const uchar2 rc = tri_rc[el];
                    Qn[el] = schur_entry_p(Pn, myld, rc.x + 2, rc.y + 2, e, i1, i2);
                }
                __syncthreads();
                pool += nodes * csz;
                { int* t = pc; pc = pn; pn = t; t = lc; lc = ln; ln = t; }
                { double* t = dc; dc = dn; dn = t; t = sc; sc = sn; sn = t; }
                nodes *= 2; --shift;
            }
            // ---- 2-mode nodes finished by single threads
            for (int nd = tid; nd < nodes; nd += THREADS) {
                if constexpr (AUG != 0) dd_add(acc, tail2_loop_p(S + pc[nd], lc[nd], dc[nd], sc[nd]));
                else dd_add(acc, tail2_p(S + pc[nd], lc[nd], dc[nd], sc[nd]));
            }
        }
    }
    __shared__ double red[(THREADS / 32) * 4];
    cdd a2;
    a2.re = acc;
    a2.im = {0.0, 0.0};
    block_reduce_store(a2, red, partials);
}

// =================================================================================================
// v4 EXPERIMENT (torontonian only, opt-in with WB200_TOR_V4=1): warp-autonomous expansion of the last 9 modes on the
// FP64 tensor cores.  Measured SLOWER than tor_kernel (2.5 ms vs 1.22 ms at 2N = 48, see tor_launch) — kept because it is
// correct, documents the rank-2 DMMA form of the Schur complement, and is the starting point for a version with more
// eliminations in flight per warp.
// =================================================================================================
// Phases A and B (shared leading modes, depth-first walk of the g group modes, all threads of the CTA) are those of
// tor_kernel.  What changes is the expansion of a prefix's 18 x 18 root: instead of a breadth-first sweep by the whole
// CTA (seven barriers per prefix, one or two matrix entries per thread and level — 21 % FP64 pipe, barrier and
// short-scoreboard stalls: profiles/r01_ncu_tor48_v3b.txt), every WARP expands a root on its own:
//   * modes 0 .. 5 of the root are walked depth-first with one buffer per depth (packed lower triangles); stepping to
//     the next of the 64 paths costs ONE elimination, and that elimination is a rank-2 complex Hermitian update
//         C <- C - i1 a a^H - i2 u u^H        (a = column 0, u = column 1 after the first pivot)
//     which in real arithmetic is a K = 4 product:  Re C -= [a_r a_i u_r u_i] diag(i1 i1 i2 i2) [a_r a_i u_r u_i]^T
//     — exactly one DMMA.8x8x4 per 8 x 8 tile for the real part and one for the imaginary part (40 flops per
//     entry in 2 tensor instructions per 64 entries instead of ~20 FP64 instructions per entry);
//   * the 64 three-mode (6 x 6) nodes of a root are collected 32 at a time and finished one per LANE in registers
//     (one 4 x 4 Schur complement and two 2-mode tails = 8 subsets per lane), so the deepest levels, where almost
//     all nodes live, run with every lane busy and no shared-memory traffic.
// No CTA barrier inside an expansion; the CTA synchronises only for the shared eliminations of phases A / B.
constexpr int T4_DC = 9;                 // modes expanded per root (the C-ABI prefix unit, as tor_kernel)
constexpr int T4_WLV = 6;                // of which walked by the warp (root dim 18 -> children 16 .. 6)
constexpr int T4_MAXW = 8;               // warps per CTA (fewer if shared memory does not allow)
constexpr int T4_LEAF = 21;              // double2 per packed 6 x 6 leaf (odd: conflict-free per quarter warp)
__host__ __device__ constexpr int t4_tri(int r, int c) { return r * (r + 1) / 2 + c; }
// per-warp area (double2 units): root | depth buffers (child dims 16, 14, 12, 10, 8) | 32 leaves | leaf (det, sgn) | fragment vectors
constexpr int T4_OFF_ROOT = 0;
constexpr int T4_OFF_DEPTH = T4_OFF_ROOT + t4_tri(18, 0);                                          // 171
constexpr int T4_OFF_LEAF = T4_OFF_DEPTH + t4_tri(16, 0) + t4_tri(14, 0) + t4_tri(12, 0) + t4_tri(10, 0) + t4_tri(8, 0);
constexpr int T4_OFF_LDS = T4_OFF_LEAF + 32 * T4_LEAF;
constexpr int T4_OFF_VEC = T4_OFF_LDS + 32;                                                         // 32 x (det, sgn)
constexpr int T4_WARP_D2 = T4_OFF_VEC + 3 * 4 * 16 / 2;                                             // 3 tables x 4 components x 16 rows of doubles

__device__ __forceinline__ void t4_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Eliminate the leading mode of the packed lower-triangular view (P, row offset k, dimension d) into the packed
// buffer Q (dimension d - 2) with one warp.  vec: 192 doubles of scratch.  Returns d1 * d2.
__device__ __forceinline__ double t4_schur(const double2* __restrict__ P, int k, int d, double2* __restrict__ Q,
                                           double* __restrict__ vec, int lane) {
    const int cd = d - 2;
    const double2 e = P[t4_tri(k + 1, k)];
    double i1, i2, d1d2;
    pivot_inverses(P[t4_tri(k, k)].x, P[t4_tri(k + 1, k + 1)].x, e, i1, i2, d1d2);
    if (lane < 16) {
        double2 a = make_double2(0.0, 0.0), u = a;
        if (lane < cd) {
            a = P[t4_tri(k + 2 + lane, k)];
            u = P[t4_tri(k + 2 + lane, k + 1)];
            const double2 q = cmulc(a, e);
            u.x -= q.x * i1; u.y -= q.y * i1;
        }
        // B operand: unscaled components; A operand (real part): -i scaled; A' operand (imaginary part)
        vec[0 * 16 + lane] = a.x;            vec[1 * 16 + lane] = a.y;           vec[2 * 16 + lane] = u.x;            vec[3 * 16 + lane] = u.y;
        vec[64 + 0 * 16 + lane] = -i1 * a.x; vec[64 + 1 * 16 + lane] = -i1 * a.y; vec[64 + 2 * 16 + lane] = -i2 * u.x; vec[64 + 3 * 16 + lane] = -i2 * u.y;
        vec[128 + 0 * 16 + lane] = -i1 * a.y; vec[128 + 1 * 16 + lane] = i1 * a.x; vec[128 + 2 * 16 + lane] = -i2 * u.y; vec[128 + 3 * 16 + lane] = i2 * u.x;
    }
    __syncwarp();
    const int g = lane >> 2, t = lane & 3;
    const int ntile = cd > 8 ? 2 : 1;
    for (int ti = 0; ti < ntile; ++ti) {
        const int row = 8 * ti + g;
        const double ar = vec[64 + t * 16 + row], ai = vec[128 + t * 16 + row];
        for (int tj = 0; tj <= ti; ++tj) {
            const int c0 = 8 * tj + 2 * t;
            const double b = vec[t * 16 + 8 * tj + g];
            const bool ok0 = row < cd && c0 <= row, ok1 = row < cd && c0 + 1 <= row;
            double2 v0 = make_double2(0.0, 0.0), v1 = v0;
            if (ok0) v0 = P[t4_tri(k + 2 + row, k + 2 + c0)];
            if (ok1) v1 = P[t4_tri(k + 2 + row, k + 2 + c0 + 1)];
            t4_dmma(v0.x, v1.x, ar, b);
            t4_dmma(v0.y, v1.y, ai, b);
            if (ok0) Q[t4_tri(row, c0)] = v0;
            if (ok1) Q[t4_tri(row, c0 + 1)] = v1;
        }
    }
    __syncwarp();
    return d1d2;
}

// One lane finishes a packed 6 x 6 (three-mode) node: 8 subsets.
__device__ __forceinline__ double t4_leaf(const double2* __restrict__ L, double det, double sgn) {
    double2 q[16];
    // exclude the leading mode: the trailing 4 x 4 block, sign flipped
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) q[r * 4 + c] = L[t4_tri(r + 2, c + 2)];
    double sum = tail2(q, 4, det, -sgn);
    // include it: both pivots fused (schur_entry on the packed triangle)
    const double2 e = L[t4_tri(1, 0)];
    double i1, i2, d1d2;
    pivot_inverses(L[0].x, L[t4_tri(1, 1)].x, e, i1, i2, d1d2);
    double2 a[4], u[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        a[r] = L[t4_tri(r + 2, 0)];
        u[r] = L[t4_tri(r + 2, 1)];
        const double2 w = cmulc(a[r], e);
        u[r].x -= w.x * i1; u[r].y -= w.y * i1;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            double2 v = q[r * 4 + c];
            { const double2 w = cmulc(a[r], a[c]); v.x -= w.x * i1; v.y -= w.y * i1; }
            { const double2 w = cmulc(u[r], u[c]); v.x -= w.x * i2; v.y -= w.y * i2; }
            q[r * 4 + c] = v;
        }
    sum += tail2(q, 4, det * d1d2, sgn);
    return sum;
}

// Expand one 18 x 18 root (packed lower triangle in W + T4_OFF_ROOT) with one warp: 512 subsets.
__device__ __forceinline__ void t4_expand(double2* __restrict__ W, double rdet, double rsgn, dd& acc, int lane) {
    double* vec = reinterpret_cast<double*>(W + T4_OFF_VEC);
    double2* lds = W + T4_OFF_LDS;
    int nbase[T4_WLV + 1], nk[T4_WLV + 1];
    double ndet[T4_WLV + 1], nsgn[T4_WLV + 1];
#pragma unroll
    for (int l = 0; l <= T4_WLV; ++l) { nbase[l] = T4_OFF_ROOT; nk[l] = 0; ndet[l] = rdet; nsgn[l] = rsgn; }
    int prev = -1, nleaf = 0;
    for (int sub = 0; sub < (1 << T4_WLV); ++sub) {
        const int first = prev < 0 ? 0 : T4_WLV - (32 - __clz(sub ^ prev));
        prev = sub;
        double2* slot = W + T4_OFF_LEAF + nleaf * T4_LEAF;
        int dbuf = T4_OFF_DEPTH;
#pragma unroll
        for (int lvl = 0; lvl < T4_WLV; ++lvl) {
            const int dim = 2 * T4_DC - 2 * lvl, cd = dim - 2;
            if (lvl >= first) {
                const bool inc = (sub >> (T4_WLV - 1 - lvl)) & 1;
                if (inc) {
                    const int dst = (lvl == T4_WLV - 1) ? (int)(slot - W) : dbuf;      // the last level writes the leaf slot itself
                    ndet[lvl + 1] = ndet[lvl] * t4_schur(W + nbase[lvl], nk[lvl], dim, W + dst, vec, lane);
                    nsgn[lvl + 1] = nsgn[lvl];
                    nbase[lvl + 1] = dst; nk[lvl + 1] = 0;
                } else {
                    ndet[lvl + 1] = ndet[lvl]; nsgn[lvl + 1] = -nsgn[lvl];
                    nbase[lvl + 1] = nbase[lvl]; nk[lvl + 1] = nk[lvl] + 2;
                }
            }
            dbuf += t4_tri(cd, 0);
        }
        // the three-mode node of this path: gather it into the leaf slot unless the last elimination wrote it there
        if (nbase[T4_WLV] != (int)(slot - W)) {
            if (lane < T4_LEAF) {
                int r = 0;                       // lane -> (r, c) of the 6 x 6 lower triangle
                while (t4_tri(r + 1, 0) <= lane) ++r;
                const int c = lane - t4_tri(r, 0), k = nk[T4_WLV];
                slot[lane] = W[nbase[T4_WLV] + t4_tri(k + r, k + c)];
            }
        }
        if (lane == 0) lds[nleaf] = make_double2(ndet[T4_WLV], nsgn[T4_WLV]);
        ++nleaf;
        if (nleaf == 32 || sub == (1 << T4_WLV) - 1) {
            __syncwarp();
            if (lane < nleaf) {
                const double2 ds = lds[lane];
                dd_add(acc, t4_leaf(W + T4_OFF_LEAF + lane * T4_LEAF, ds.x, ds.y));
            }
            __syncwarp();
            nleaf = 0;
        }
    }
}

// shared-memory plan of tor4_kernel (double2 units): T | depth buffers of the group walk | per-warp areas | root (det, sgn)
__host__ __device__ inline size_t tor4_smem_plan(int N, int g, int warps, int* off_depth, int* off_warp) {
    const int n2 = 2 * N, dg = 2 * (T4_DC + g);
    int off = n2 * tor_ld(n2);
    *off_depth = off;
    for (int k = 1; k <= g; ++k) off += (dg - 2 * k) * tor_ld(dg - 2 * k);
    *off_warp = off;
    off += warps * T4_WARP_D2;
    return (size_t)off * sizeof(double2) + 2 * T4_MAXW * sizeof(double);
}

__global__ void __launch_bounds__(32 * T4_MAXW) tor4_kernel(TorParams p, int off_warp, double* __restrict__ partials) {
    extern __shared__ __align__(16) double smem_tor[];
    const int N = p.N, n2 = 2 * N, g = p.g, P = p.P;
    const int dg = 2 * (T4_DC + g), ldT = tor_ld(n2);
    const int threads = blockDim.x, nwarps = threads >> 5;
    double2* S = reinterpret_cast<double2*>(smem_tor);
    double* rootds = reinterpret_cast<double*>(S + off_warp + nwarps * T4_WARP_D2);     // (det, sgn) of the collected roots
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double2* W = S + off_warp + warp * T4_WARP_D2;

    // eliminate the leading mode of view (V, ld, dim) into (Q, ldq) with all threads (eliminate_into for a runtime CTA size)
    auto eliminate = [&](const double2* V, int ld, int dim, double2* Q, int ldq) -> double {
        __syncthreads();
        const double2 e = V[ld];
        double i1, i2, d1d2;
        pivot_inverses(V[0].x, V[ld + 1].x, e, i1, i2, d1d2);
        const int cd = dim - 2, tri = cd * (cd + 1) / 2;
        for (int el = tid; el < tri; el += threads) {
            int r, c;
            tri_decode(el, r, c);
            Q[r * ldq + c] = schur_entry(V, ld, r + 2, c + 2, e, i1, i2);
        }
        __syncthreads();
        return d1d2;
    };

    dd acc = {0.0, 0.0};
    const uint64_t ngroups_first = p.p0 >> g, ngroups_last = (p.p1 + (1ull << g) - 1) >> g;
    for (uint64_t grp = ngroups_first + blockIdx.x; grp < ngroups_last; grp += gridDim.x) {
        // ---- phase A: common leading modes 0 .. P-g-1 (bits of grp, most significant = mode 0), in place on T
        __syncthreads();
        for (int idx = tid; idx < n2 * n2; idx += threads) S[(idx / n2) * ldT + idx % n2] = static_cast<const double2*>(p.B)[idx];
        __syncthreads();
        double det0 = 1.0, sgn0 = 1.0;
        const int lead = P - g;
        for (int i = 0; i < lead; ++i) {
            const bool inc = (grp >> (lead - 1 - i)) & 1ull;
            double2* V = S + 2 * i * (ldT + 1);
            if (inc) det0 *= eliminate(V, ldT, n2 - 2 * i, V + 2 * (ldT + 1), ldT);
            else sgn0 = -sgn0;
        }
        // ---- phase B: the 2^g prefixes of this group, depth-first; roots are handed to the warps nwarps at a time
        int nptr[TOR_MAXG + 1], nld[TOR_MAXG + 1];
        double ndet[TOR_MAXG + 1], nsgn[TOR_MAXG + 1];
        nptr[0] = 2 * lead * (ldT + 1); nld[0] = ldT; ndet[0] = det0; nsgn[0] = sgn0;
#pragma unroll
        for (int k = 1; k <= TOR_MAXG; ++k) { nptr[k] = 0; nld[k] = 1; ndet[k] = 1.0; nsgn[k] = 1.0; }
        int cur = -1, nroots = 0;
        for (int sub = 0; sub < (1 << g); ++sub) {
            const uint64_t pfx = (grp << g) + sub;
            const bool live = pfx >= p.p0 && pfx < p.p1;
            if (live) {
                const int first = cur < 0 ? 0 : g - (32 - __clz(sub ^ cur));
                cur = sub;
                int dbuf = p.off_depth;
#pragma unroll
                for (int lvl = 0; lvl < TOR_MAXG; ++lvl) {
                    if (lvl < g) {
                        const int dim = dg - 2 * lvl, cdim = dim - 2, cld = tor_ld(cdim);
                        if (lvl >= first) {
                            const bool inc = (sub >> (g - 1 - lvl)) & 1;
                            if (inc) {
                                ndet[lvl + 1] = ndet[lvl] * eliminate(S + nptr[lvl], nld[lvl], dim, S + dbuf, cld);
                                nsgn[lvl + 1] = nsgn[lvl];
                                nptr[lvl + 1] = dbuf; nld[lvl + 1] = cld;
                            } else {
                                ndet[lvl + 1] = ndet[lvl]; nsgn[lvl + 1] = -nsgn[lvl];
                                nptr[lvl + 1] = nptr[lvl] + 2 * (nld[lvl] + 1); nld[lvl + 1] = nld[lvl];
                            }
                        }
                        dbuf += cdim * cld;
                    }
                }
                int rptr = nptr[0], rld = nld[0];
                double rdet = ndet[0], rsgn = nsgn[0];
#pragma unroll
                for (int k = 1; k <= TOR_MAXG; ++k)
                    if (k == g) { rptr = nptr[k]; rld = nld[k]; rdet = ndet[k]; rsgn = nsgn[k]; }
                // warp `nroots` copies the root (a view of CTA-level storage, stable until the next elimination's barrier)
                if (warp == nroots) {
                    const double2* R = S + rptr;
                    double2* dst = W + T4_OFF_ROOT;
                    for (int el = lane; el < t4_tri(2 * T4_DC, 0); el += 32) {
                        int r, c;
                        tri_decode(el, r, c);
                        dst[el] = R[r * rld + c];
                    }
                    if (lane == 0) { rootds[2 * warp] = rdet; rootds[2 * warp + 1] = rsgn; }
                    __syncwarp();
                }
                ++nroots;
            }
            if (nroots == nwarps || (sub == (1 << g) - 1 && nroots > 0)) {
                if (warp < nroots) t4_expand(W, rootds[2 * warp], rootds[2 * warp + 1], acc, lane);
                nroots = 0;
                __syncthreads();     // the next root copies and eliminations may overwrite what the expansions read
            }
        }
    }
    __shared__ double red[T4_MAXW * 4];
    cdd a2;
    a2.re = acc;
    a2.im = {0.0, 0.0};
    block_reduce_store(a2, red, partials);
}

// DC modes expanded breadth-first, 2^g prefixes per CTA group; g shrinks if shared memory does not allow 5.
static void tor_shape(int N, int aug, int* P, int* g, int* DC) {
    *DC = N < TOR_DC ? N : TOR_DC;
    *P = N - *DC;
    *g = *P < TOR_G ? *P : TOR_G;
    int a, b, c;
    while (*g > 0 && tor_smem_plan_p(N, aug, *g, *DC, &a, &b, &c, sizeof(double2)) > (size_t)WB_TOR_SMEM_KB * 1024) --*g;
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
    tor_shape(n_modes, 0, &P, &g, &DC);
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

// shared launcher: dGamma == nullptr -> torontonian, else loop torontonian; real_O: dO is a REAL symmetric 2N x 2N matrix
// (torontonian only) and the tree runs in real arithmetic
static int tor_launch(const double* dO, const double* dGamma, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                      void* d_workspace, size_t workspace_bytes, void* stream, bool real_O = false) {
    if (!dO || !d_out4 || !d_workspace) { set_error("tor: null pointer"); return WB200_EINVAL; }
    uint64_t total = 0;
    int rc = wb200_tor_num_prefixes(n_modes, &total);
    if (rc) return rc;
    if (p0 > p1 || p1 > total) { set_error("tor: bad prefix range"); return WB200_EINVAL; }
    if (workspace_bytes < wb200_tor_workspace_bytes(n_modes)) { set_error("tor: workspace too small"); return WB200_EINVAL; }
    const int N = n_modes, aug = dGamma ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    TorParams p;
    tor_shape(N, aug, &p.P, &p.g, &p.DC);
    p.N = N; p.p0 = p0; p.p1 = p1;
    int dev = 0, sms = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (device_sm_count(dev, &sms)) return WB200_ECUDA;
    double2* dB = reinterpret_cast<double2*>(d_workspace);
    double* dpart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_workspace) + tor_ws_partials_offset(N));
    p.B = dB;
    const char* ev4 = getenv("WB200_TOR_V4");
    if (!aug && !real_O && p.DC == T4_DC && ev4 && atoi(ev4)) {
        // v4 EXPERIMENT (opt-in, WB200_TOR_V4=1): warp-autonomous tensor-core expansion (tor4_kernel).  Correct (same
        // parity tests pass), but 2.5 ms at 2N = 48 against 1.22 ms for the breadth-first kernel below: the depth-first
        // chain of a warp is one long latency chain (pivot loads -> FP64 division -> fragment vectors -> DMMA -> store,
        // ~2000 cycles per elimination at 0.13 IPC) and shared memory only admits 7 such warps per SM
        // (profiles/r02_ncu_tor48_v4.txt, profiles/r02_tor4_shapes.txt).  g group modes and as many warps as shared
        // memory allows next to the CTA-level buffers; env WB200_TOR4_G / WB200_TOR4_W override.
        const char* eg = getenv("WB200_TOR4_G");
        const char* ew = getenv("WB200_TOR4_W");
        int g = p.P < 4 ? p.P : 4;
        if (eg) { g = atoi(eg); if (g > p.P) g = p.P; if (g > TOR_MAXG) g = TOR_MAXG; if (g < 0) g = 0; }
        while (g > 0 && ((p1 - p0) >> g) < 2ull * (uint64_t)sms) --g;     // thin shards: every SM gets a group
        int warps = ew ? atoi(ew) : T4_MAXW;
        if (warps > T4_MAXW) warps = T4_MAXW;
        if (warps < 1) warps = 1;
        int off_warp = 0;
        size_t shm = 0;
        for (;;) {
            shm = tor4_smem_plan(N, g, warps, &p.off_depth, &off_warp);
            if (shm <= (size_t)226 * 1024) break;
            if (warps > 4) --warps; else if (g > 0) --g; else break;
        }
        if (shm > (size_t)226 * 1024) { set_error("tor: %d modes need %zu bytes of shared memory", N, shm); return WB200_ENOSUP; }
        p.g = g;
        WB_CUDA(cudaFuncSetAttribute(tor4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        const uint64_t groups = ((p1 + (1ull << g) - 1) >> g) - (p0 >> g);
        int grid = (int)(groups < (uint64_t)sms ? (groups ? groups : 1) : (uint64_t)sms);
        tor_prep_kernel<<<8, 256, 0, st>>>(reinterpret_cast<const double2*>(dO), nullptr, N, dB);
        tor4_kernel<<<grid, 32 * warps, shm, st>>>(p, off_warp, dpart);
        final_reduce_kernel<<<1, 32, 0, st>>>(dpart, grid, d_out4);
        WB_CUDA(cudaGetLastError());
        return WB200_OK;
    }
    // small problems (or thin multi-GPU shards): fewer prefixes per CTA so that every SM gets a group
    while (p.g > 0 && ((p1 - p0) >> p.g) < 2ull * (uint64_t)sms) --p.g;
    if (real_O) {
        const size_t shm = tor_smem_plan_p(N, 0, p.g, p.DC, &p.off_depth, &p.off_pool, &p.off_desc, sizeof(double));
        if (shm > 226 * 1024) { set_error("tor: %d modes need %zu bytes of shared memory", N, shm); return WB200_ENOSUP; }
        auto kern = tor_kernel<0, TOR_THREADS, double>;
        WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        const uint64_t groups = ((p1 + (1ull << p.g) - 1) >> p.g) - (p0 >> p.g);
        int occ = 1;
        WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TOR_THREADS, shm));
        if (occ < 1) occ = 1;
        const uint64_t slots = (uint64_t)sms * occ;
        int grid = (int)(groups < slots ? (groups ? groups : 1) : slots);
        if (grid > TOR_MAX_GRID) grid = TOR_MAX_GRID;
        tor_prep_real_kernel<<<8, 256, 0, st>>>(dO, N, reinterpret_cast<double*>(dB));
        kern<<<grid, TOR_THREADS, shm, st>>>(p, dpart);
        final_reduce_kernel<<<1, 32, 0, st>>>(dpart, grid, d_out4);
        WB_CUDA(cudaGetLastError());
        return WB200_OK;
    }
    const size_t shm = tor_smem_plan_p(N, aug, p.g, p.DC, &p.off_depth, &p.off_pool, &p.off_desc, sizeof(double2));
    if (shm > 226 * 1024) { set_error("tor: %d modes need %zu bytes of shared memory", N, shm); return WB200_ENOSUP; }
    if (aug) WB_CUDA(cudaFuncSetAttribute(tor_kernel<1, TOR_THREADS_LOOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    else WB_CUDA(cudaFuncSetAttribute(tor_kernel<0, TOR_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    const uint64_t groups = ((p1 + (1ull << p.g) - 1) >> p.g) - (p0 >> p.g);
    int occ = 1;
    if (aug) WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tor_kernel<1, TOR_THREADS_LOOP>, TOR_THREADS_LOOP, shm));
    else WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tor_kernel<0, TOR_THREADS>, TOR_THREADS, shm));
    if (occ < 1) occ = 1;
    const uint64_t slots = (uint64_t)sms * occ;
    int grid = (int)(groups < slots ? (groups ? groups : 1) : slots);
    if (grid > TOR_MAX_GRID) grid = TOR_MAX_GRID;
    tor_prep_kernel<<<8, 256, 0, st>>>(reinterpret_cast<const double2*>(dO), reinterpret_cast<const double2*>(dGamma), N, dB);
    if (aug) tor_kernel<1, TOR_THREADS_LOOP><<<grid, TOR_THREADS_LOOP, shm, st>>>(p, dpart);
    else tor_kernel<0, TOR_THREADS><<<grid, TOR_THREADS, shm, st>>>(p, dpart);
    final_reduce_kernel<<<1, 32, 0, st>>>(dpart, grid, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

extern "C" int wb200_tor_dev(const double* dO, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
    return tor_launch(dO, nullptr, n_modes, p0, p1, d_out4, d_workspace, workspace_bytes, stream);
}

extern "C" int wb200_tor_f64_dev(const double* dO, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                                 void* d_workspace, size_t workspace_bytes, void* stream) {
    return tor_launch(dO, nullptr, n_modes, p0, p1, d_out4, d_workspace, workspace_bytes, stream, true);
}

extern "C" int wb200_ltor_dev(const double* dO, const double* dGamma, int n_modes, uint64_t p0, uint64_t p1,
                              double* d_out4, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!dGamma) { set_error("ltor: null gamma"); return WB200_EINVAL; }
    return tor_launch(dO, dGamma, n_modes, p0, p1, d_out4, d_workspace, workspace_bytes, stream);
}

static int tor_host_impl(int device, const double* O, const double* gamma, int n_modes, uint64_t p0, uint64_t p1,
                         double out2[2], double* kernel_ms, bool real_O = false) {
    if (!O || !out2) { set_error("tor: null pointer"); return WB200_EINVAL; }
    uint64_t total = 0;
    int rc = wb200_tor_num_prefixes(n_modes, &total);
    if (rc) return rc;
    const int n2 = 2 * n_modes;
    WB_CUDA(cudaSetDevice(device));
    DevBufT dO, dG, dws, dout;
    const size_t wsb = wb200_tor_workspace_bytes(n_modes);
    const size_t obytes = sizeof(double) * (real_O ? 1 : 2) * n2 * n2;
    WB_POOL(pool_alloc(&dO.p, obytes));
    WB_POOL(pool_alloc(&dws.p, wsb));
    WB_POOL(pool_alloc(&dout.p, sizeof(double) * 4));
    WB_CUDA(cudaMemcpy(dO.p, O, obytes, cudaMemcpyHostToDevice));
    if (gamma) {
        WB_POOL(pool_alloc(&dG.p, sizeof(double) * 2 * n2));
        WB_CUDA(cudaMemcpy(dG.p, gamma, sizeof(double) * 2 * n2, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) {        // events only when the caller wants the kernel time (the end-to-end path does not)
        WB_CUDA(cudaEventCreate(&e0));
        WB_CUDA(cudaEventCreate(&e1));
        WB_CUDA(cudaEventRecord(e0, 0));
    }
    rc = tor_launch((const double*)dO.p, (const double*)dG.p, n_modes, p0, p1, (double*)dout.p, dws.p, wsb, nullptr, real_O);
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
    double o4[4];
    WB_CUDA(cudaMemcpy(o4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    out2[0] = o4[0]; out2[1] = o4[1];
    return WB200_OK;
}

extern "C" int wb200_tor_host(int device, const double* O, int n_modes, uint64_t p0, uint64_t p1, double out2[2],
                              double* kernel_ms) {
    return tor_host_impl(device, O, nullptr, n_modes, p0, p1, out2, kernel_ms);
}

extern "C" int wb200_tor_f64_host(int device, const double* O, int n_modes, uint64_t p0, uint64_t p1, double out2[2],
                                  double* kernel_ms) {
    return tor_host_impl(device, O, nullptr, n_modes, p0, p1, out2, kernel_ms, true);
}

extern "C" int wb200_ltor_host(int device, const double* O, const double* gamma, int n_modes, uint64_t p0, uint64_t p1,
                               double out2[2], double* kernel_ms) {
    if (!gamma) { set_error("ltor: null gamma"); return WB200_EINVAL; }
    return tor_host_impl(device, O, gamma, n_modes, p0, p1, out2, kernel_ms);
}
