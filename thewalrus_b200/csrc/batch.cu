// Batched loop-hafnian front ends (one warp per subset, repeated edges, Glynn or inclusion/exclusion):
//
//  * lhaf_patterns: (A, gamma, rpt[B][nv]) -> loop hafnian of every repetition pattern.  Replaces the
//    Python loop of probabilities() over density_matrix_element (thewalrus/quantum/fock_tensors.py:191-232,
//    412-421), i.e. per pattern matched_reps (thewalrus/_hafnian.py:80-159) + _calc_loop_hafnian (:512-577).
//    (pattern, subset) pairs are flattened into one index space: a prep kernel pairs the vertices of every
//    pattern and counts its subsets, a scan turns the per-pattern chunk counts into offsets, persistent
//    warps pull fixed-size chunks from an atomic counter and write one compensated partial per chunk, and a
//    final kernel adds each pattern's chunks in index order (deterministic: no floating-point atomics).
//  * lhaf_batch: _calc_loop_hafnian_batch_even / _odd (thewalrus/loop_hafnian_batch.py:51-208): one sweep
//    over the subsets of (batch edge + fixed edges); every subset contributes to all photon numbers
//    N_det >= 2 kept_0 of the batch mode.
//
// Per subset (all in the warp's slice of shared memory): reduced matrix M = AX_S (get_submatrices,
// _hafnian.py:315-356), power traces by a product chain with the pairing
// tr(M^(a+b)) = sum_rc (M^a)[r,c] (M^b)[c,r] — continued past the matrix size, where the reference switches
// to La Budde + Newton (thewalrus/charpoly.py:319-326; same numbers) — loop terms XD M^(t-1) D and
// oddVX M^(t-1) D from one mat-vec chain, and the series c_t = (1/t) sum_i i a_i c_(t-i) of f_loop /
// f_loop_odd (_hafnian.py:212-285).
#include <stdlib.h>
#include "common.cuh"

namespace wb {

constexpr int BW_EMAX = 32;      // max edges per problem
constexpr int BW_NVMAX = 64;     // max vertices of the big matrix
constexpr int BW_WARPS = 8;      // warps per CTA
#ifndef WB_BW_CHUNK
#define WB_BW_CHUNK 16
#endif
constexpr int BW_CHUNK = WB_BW_CHUNK;   // subsets per work item (patterns kernel)
constexpr int BW_MAX_ORDER = 200;

__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 warp_sum2(double2 v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { v.x += shfl_xor_d(v.x, off); v.y += shfl_xor_d(v.y, off); }
    return v;
}
__device__ __forceinline__ double binom_d(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    if (k > n - k) k = n - k;
    double b = 1.0;
    for (int q = 0; q < k; ++q) b = b * (double)(n - q) / (double)(q + 1);
    return rint(b);
}

// per-warp shared-memory workspace
struct WarpWs {
    double2 *M, *P, *Pn;              // smax * smax each
    double2 *vXD, *vD, *vD2, *vOV, *vOV0;  // smax each
    double2 *ptr, *lv, *ov, *ov0;     // T + 2 each
    double2 *fac, *cs0, *cs1, *cs2;   // O + 2 each
    int* rows;                        // 2 * BW_EMAX
    double* delta;                    // 2 * BW_EMAX
    int* kept;                        // BW_EMAX
    unsigned char* eu; unsigned char* ev; unsigned short* er;  // edge lists (BW_EMAX each)
};
__host__ __device__ inline size_t warp_ws_bytes(int smax, int T, int O) {
    size_t b = sizeof(double2) * ((size_t)3 * smax * smax + 5 * (size_t)smax + 4 * (size_t)(T + 2) + 4 * (size_t)(O + 2));
    b += sizeof(int) * 2 * BW_EMAX + sizeof(double) * 2 * BW_EMAX + sizeof(int) * BW_EMAX + 2 * BW_EMAX + 2 * BW_EMAX;
    return (b + 15) & ~(size_t)15;
}
__device__ inline WarpWs carve(unsigned char* base, int smax, int T, int O) {
    WarpWs w;
    double2* p = reinterpret_cast<double2*>(base);
    w.M = p; p += smax * smax; w.P = p; p += smax * smax; w.Pn = p; p += smax * smax;
    w.vXD = p; p += smax; w.vD = p; p += smax; w.vD2 = p; p += smax; w.vOV = p; p += smax; w.vOV0 = p; p += smax;
    w.ptr = p; p += T + 2; w.lv = p; p += T + 2; w.ov = p; p += T + 2; w.ov0 = p; p += T + 2;
    w.fac = p; p += O + 2; w.cs0 = p; p += O + 2; w.cs1 = p; p += O + 2; w.cs2 = p; p += O + 2;
    double* d = reinterpret_cast<double*>(p);
    w.delta = d; d += 2 * BW_EMAX;
    int* ip = reinterpret_cast<int*>(d);
    w.rows = ip; ip += 2 * BW_EMAX; w.kept = ip; ip += BW_EMAX;
    unsigned short* sp = reinterpret_cast<unsigned short*>(ip);
    w.er = sp; sp += BW_EMAX;
    unsigned char* cp = reinterpret_cast<unsigned char*>(sp);
    w.eu = cp; cp += BW_EMAX; w.ev = cp;
    return w;
}

// Decode subset j (mixed radix, most significant digit first: find_kept_edges, _hafnian.py:162-180), build the
// list of kept rows.  Returns k (edges with delta != 0); *esum = sum kept, *weight = prod_{i >= wstart} C(r_i, kept_i),
// *d0zero = (delta_0 == 0).  Executed by lane 0; results broadcast by the caller through shared memory.
__device__ inline int decode_subset(const WarpWs& w, int E, int glynn, unsigned long long j, int wstart, int* esum,
                                    double* weight, int* d0zero) {
    unsigned long long num = j;
    for (int i = E - 1; i >= 0; --i) {
        const unsigned long long base = (unsigned long long)w.er[i] + 1ull;
        w.kept[i] = (int)(num % base);
        num /= base;
    }
    int k = 0, es = 0;
    double wt = 1.0;
    *d0zero = 0;
    for (int i = 0; i < E; ++i) {
        const int r = w.er[i], kp = w.kept[i];
        es += kp;
        if (i >= wstart) wt *= binom_d(r, kp);
        const int d = glynn ? 2 * kp - r : kp;
        if (i == 0) *d0zero = (d == 0);
        if (d != 0) { w.rows[k] = i; w.delta[k] = (double)d; ++k; }
    }
    for (int a = 0; a < k; ++a) {
        const int e = w.rows[a];
        w.rows[a] = w.eu[e];
        w.rows[k + a] = w.ev[e];
        w.delta[k + a] = w.delta[a];
    }
    *esum = es; *weight = wt;
    return k;
}

// Reduced matrix M = AX_S, odd-row vectors and power traces ptr[1..T] of subset (rows, delta) — everything that
// does not depend on the loop vector D.  A: nv x nv (row stride lda); odd rows are rows of A (index or -1).
__device__ inline void subset_setup(const WarpWs& w, const double2* __restrict__ A, int lda, int odd_row, int odd0_row,
                                    int k, int T, int lane) {
    const int s = 2 * k;
    for (int idx = lane; idx < s * s; idx += 32) {
        const int r = idx / s, c = idx - r * s;
        const int sc = c < k ? c + k : c - k;
        const double2 a = __ldg(A + (size_t)w.rows[r] * lda + w.rows[sc]);
        const double d = w.delta[c];
        const double2 v = make_double2(a.x * d, a.y * d);
        w.M[idx] = v;
        w.P[idx] = v;
    }
    for (int c = lane; c < s; c += 32) {
        const int sc = c < k ? c + k : c - k;
        const double d = w.delta[c];
        if (odd_row >= 0) {
            const double2 ov = __ldg(A + (size_t)odd_row * lda + w.rows[sc]);
            w.vOV[c] = make_double2(ov.x * d, ov.y * d);
        }
        if (odd0_row >= 0) {
            const double2 ov = __ldg(A + (size_t)odd0_row * lda + w.rows[sc]);
            w.vOV0[c] = make_double2(ov.x * d, ov.y * d);
        }
    }
    __syncwarp();
    // ---- power traces
    double2 t1 = make_double2(0.0, 0.0);
    for (int r = lane; r < s; r += 32) { t1.x += w.M[r * s + r].x; t1.y += w.M[r * s + r].y; }
    t1 = warp_sum2(t1);
    if (lane == 0) { w.ptr[0] = make_double2((double)s, 0.0); if (T >= 1) w.ptr[1] = t1; }
    double2* Pc = w.P;
    double2* Pnx = w.Pn;
    for (int t = 1; 2 * t <= T; ++t) {
        double2 e = make_double2(0.0, 0.0);
        for (int idx = lane; idx < s * s; idx += 32) {
            const int r = idx / s, c = idx - r * s;
            cfma(e, Pc[idx], Pc[c * s + r]);
        }
        e = warp_sum2(e);
        if (lane == 0) w.ptr[2 * t] = e;
        if (2 * t + 1 > T) break;
        // Pnx = Pc * M with a 2 x 2 register tile per lane (s is even): four shared-memory loads feed four complex
        // multiply-adds, half the LDS traffic and a quarter of the index arithmetic of one output per lane
        {
            const int h = s >> 1;
            for (int t2 = lane; t2 < h * h; t2 += 32) {
                const int tr = t2 / h, tc = t2 - tr * h;
                const double2* p0 = Pc + 2 * tr * s;
                const double2* p1 = p0 + s;
                const double2* mc = w.M + 2 * tc;
                double2 a00 = make_double2(0.0, 0.0), a01 = a00, a10 = a00, a11 = a00;
                for (int q = 0; q < s; ++q) {
                    const double2 x0 = p0[q], x1 = p1[q];
                    const double2 y0 = mc[q * s], y1 = mc[q * s + 1];
                    cfma(a00, x0, y0); cfma(a01, x0, y1);
                    cfma(a10, x1, y0); cfma(a11, x1, y1);
                }
                double2* o = Pnx + 2 * tr * s + 2 * tc;
                o[0] = a00; o[1] = a01; o[s] = a10; o[s + 1] = a11;
            }
        }
        __syncwarp();
        double2 o = make_double2(0.0, 0.0);
        for (int idx = lane; idx < s * s; idx += 32) {
            const int r = idx / s, c = idx - r * s;
            cfma(o, Pnx[idx], Pc[c * s + r]);
        }
        o = warp_sum2(o);
        if (lane == 0) w.ptr[2 * t + 1] = o;
        double2* tmp = Pc; Pc = Pnx; Pnx = tmp;
        __syncwarp();
    }
    __syncwarp();
}

// Loop terms of one loop vector D (nv complex, or null): lv[t] = XD M^(t-1) D, ov[t] = oddVX M^(t-1) D and
// ov0[t] for the second odd row, t = 1..T, from one mat-vec chain on the M left by subset_setup.
// (The montrealer uses the general form: left vector Dl gathered through the pair swap, right vector Dr gathered
// with or without it.)
__device__ inline void subset_loops_lr(const WarpWs& w, const double2* __restrict__ Dl, const double2* __restrict__ D,
                                       bool swap_right, bool has_odd, bool has_odd0, int k, int T, int lane) {
    const int s = 2 * k;
    if (D == nullptr) {
        for (int t = lane + 1; t <= T; t += 32) { w.lv[t] = make_double2(0.0, 0.0); w.ov[t] = w.lv[t]; w.ov0[t] = w.lv[t]; }
        __syncwarp();
        return;
    }
    for (int c = lane; c < s; c += 32) {
        const int sc = c < k ? c + k : c - k;
        const double d = w.delta[c];
        const double2 dv = __ldg(Dl + w.rows[sc]);
        w.vXD[c] = make_double2(dv.x * d, dv.y * d);
        w.vD[c] = __ldg(D + w.rows[swap_right ? sc : c]);
    }
    __syncwarp();
    double2* v = w.vD;
    double2* v2 = w.vD2;
    for (int t = 1; t <= T; ++t) {
        double2 l = make_double2(0.0, 0.0), o = make_double2(0.0, 0.0), o0 = make_double2(0.0, 0.0);
        for (int c = lane; c < s; c += 32) {
            const double2 dv = v[c];
            cfma(l, w.vXD[c], dv);
            if (has_odd) cfma(o, w.vOV[c], dv);
            if (has_odd0) cfma(o0, w.vOV0[c], dv);
        }
        l = warp_sum2(l);
        if (has_odd) o = warp_sum2(o);
        if (has_odd0) o0 = warp_sum2(o0);
        if (lane == 0) { w.lv[t] = l; w.ov[t] = o; w.ov0[t] = o0; }
        if (t < T) {
            for (int r = lane; r < s; r += 32) {
                double2 a = make_double2(0.0, 0.0);
                for (int q = 0; q < s; ++q) cfma(a, w.M[r * s + q], v[q]);
                v2[r] = a;
            }
            __syncwarp();
            double2* tmp = v; v = v2; v2 = tmp;
        }
    }
    // leave the chain buffers where subset_loops expects them next time (vD is rewritten from D on entry)
    __syncwarp();
}

__device__ inline void subset_loops(const WarpWs& w, const double2* __restrict__ D, bool has_odd, bool has_odd0, int k,
                                    int T, int lane) {
    subset_loops_lr(w, D, D, false, has_odd, has_odd0, k, T, lane);
}

__device__ inline void subset_traces(const WarpWs& w, const double2* __restrict__ A, int lda, const double2* __restrict__ D,
                                     int odd_row, int odd0_row, int k, int T, int lane) {
    subset_setup(w, A, lda, odd_row, odd0_row, k, T, lane);
    subset_loops(w, D, odd_row >= 0, odd0_row >= 0, k, T, lane);
}

// fac[i] = i * a_i for the even series (f / f_loop): a_i = p_i/(2i) + l_i/2, i = 1..order
__device__ inline void fac_even(const WarpWs& w, int order, int lane) {
    for (int i = lane + 1; i <= order; i += 32)
        w.fac[i] = make_double2(0.5 * w.ptr[i].x + 0.5 * i * w.lv[i].x, 0.5 * w.ptr[i].y + 0.5 * i * w.lv[i].y);
    __syncwarp();
}
// odd series (f_loop_odd): a_1 = oddloop, a_2t = p_t/(2t) + l_t/2, a_(2t+1) = o_t
__device__ inline void fac_odd(const WarpWs& w, int order, double2 oddloop, const double2* ov, int lane) {
    for (int i = lane + 1; i <= order; i += 32) {
        double2 f;
        if (i == 1) f = oddloop;
        else if ((i & 1) == 0) { const int t = i >> 1; f = make_double2(w.ptr[t].x + 0.5 * i * w.lv[t].x, w.ptr[t].y + 0.5 * i * w.lv[t].y); }
        else { const int t = i >> 1; f = make_double2(i * ov[t].x, i * ov[t].y); }
        w.fac[i] = f;
    }
    __syncwarp();
}
// cs[t] = (1/t) sum_{i=1..t} fac[i] cs[t-i]
__device__ inline void exp_series(const WarpWs& w, double2* cs, int order, int lane) {
    if (lane == 0) cs[0] = make_double2(1.0, 0.0);
    __syncwarp();
    for (int t = 1; t <= order; ++t) {
        double2 a = make_double2(0.0, 0.0);
        for (int i = 1 + lane; i <= t; i += 32) cfma(a, w.fac[i], cs[t - i]);
        a = warp_sum2(a);
        if (lane == 0) cs[t] = make_double2(a.x / t, a.y / t);
        __syncwarp();
    }
}

// =================================================================================================
// patterns
// =================================================================================================
struct __align__(16) PatDesc {
    unsigned long long steps;
    unsigned int nchunks;
    short E, N, odd, kind;  // kind: 0 subset sum, 1 -> 1.0, 2 -> 0.0, 3 -> D[odd]
    short cls;              // tile-shape class of the DMMA path (pat_dmma.cuh), PD_NCLS = warp-per-subset DFMA kernel
    unsigned short r[BW_EMAX];
    unsigned char u[BW_EMAX], v[BW_EMAX];
};

struct PatMeta {  // maxima over the patterns of the fallback class + error flag, filled by the prep kernel
    int maxE, maxN, anyOdd, err;
};

constexpr int PAT_NCLS = 8;      // 7 DMMA tile classes + the fallback class (== PD_NCLS + 1, checked below)
__host__ __device__ inline int pat_class_of(int E, int N, int odd);

__global__ void pat_prep_kernel(const int32_t* __restrict__ rpt, long long B, int nv, int loops, int glynn, int force_fallback,
                                const int32_t* __restrict__ aidx, int n_A, const int32_t* __restrict__ gidx, int n_gamma,
                                PatDesc* __restrict__ desc, unsigned int* __restrict__ nchunks,
                                unsigned char* __restrict__ cls_out, PatMeta* meta) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    if ((aidx && (aidx[p] < 0 || aidx[p] >= n_A)) || (gidx && loops && (gidx[p] < 0 || gidx[p] >= n_gamma))) atomicExch(&meta->err, 3);
    int cnt[BW_NVMAX];
    int N = 0, npool = 0;
    bool bad = false;
    for (int i = 0; i < nv; ++i) {
        const int r = rpt[p * nv + i];
        if (r < 0 || r > 65535) bad = true;
        cnt[i] = r < 0 ? 0 : r;
        N += cnt[i];
        npool += cnt[i] > 0;
    }
    PatDesc d;
    d.steps = 0; d.nchunks = 0; d.E = 0; d.N = (short)N; d.odd = -1; d.kind = 0; d.cls = PAT_NCLS - 1;
    cls_out[p] = (unsigned char)(PAT_NCLS - 1);
    if (bad || N > 32000) { atomicExch(&meta->err, 1); d.kind = 2; desc[p] = d; nchunks[p] = 0; return; }
    if (N == 0) d.kind = 1;
    else if (!loops && (N & 1)) d.kind = 2;
    else if (loops && N == 1) {
        d.kind = 3;
        for (int i = 0; i < nv; ++i) if (cnt[i] == 1) d.odd = (short)i;
    }
    if (d.kind != 0) { desc[p] = d; nchunks[p] = 0; return; }
    // greedy pairing (matched_reps, _hafnian.py:97-159): largest (count, index) first
    int E = 0;
    while (npool >= 1) {
        int i0 = -1, i1 = -1;
        for (int i = nv - 1; i >= 0; --i) {  // descending index so ties pick the larger index first
            if (cnt[i] <= 0) continue;
            if (i0 < 0 || cnt[i] > cnt[i0]) { i1 = i0; i0 = i; }
            else if (i1 < 0 || cnt[i] > cnt[i1]) i1 = i;
        }
        const int r0 = cnt[i0];
        if (npool == 1 && r0 <= 1) break;
        if (E >= BW_EMAX) { bad = true; break; }
        if (npool == 1 || r0 > 2 * cnt[i1]) {
            d.u[E] = (unsigned char)i0; d.v[E] = (unsigned char)i0; d.r[E] = (unsigned short)(r0 / 2);
            if (r0 & 1) cnt[i0] = 1; else { cnt[i0] = 0; --npool; }
        } else {
            const int r1 = cnt[i1];
            d.u[E] = (unsigned char)i0; d.v[E] = (unsigned char)i1; d.r[E] = (unsigned short)r1;
            cnt[i1] = 0; --npool;
            if (r0 > r1) cnt[i0] = r0 - r1; else { cnt[i0] = 0; --npool; }
        }
        ++E;
    }
    if (npool > 1) bad = true;
    if (npool == 1) for (int i = 0; i < nv; ++i) if (cnt[i] > 0) d.odd = (short)i;
    for (int e = E; e < BW_EMAX; ++e) { d.u[e] = 0; d.v[e] = 0; d.r[e] = 0; }
    // steps (_hafnian.py:432-435, 535-538)
    double stepsd = 1.0;
    unsigned long long steps = 1;
    for (int e = 0; e < E; ++e) {
        unsigned long long f = (unsigned long long)d.r[e] + 1ull;
        if (e == 0 && glynn && d.odd < 0) f = ((unsigned long long)d.r[0] + 2ull) / 2ull;
        steps *= f; stepsd *= (double)f;
    }
    const int order = d.odd >= 0 ? N : N / 2;
    if (bad || stepsd > 1e12 || order > BW_MAX_ORDER || (d.odd >= 0 && !loops)) {
        atomicExch(&meta->err, 2); d.kind = 2; desc[p] = d; nchunks[p] = 0; return;
    }
    d.E = (short)E; d.steps = steps;
    d.nchunks = (unsigned int)((steps + BW_CHUNK - 1) / BW_CHUNK);
    d.cls = (short)(force_fallback ? PAT_NCLS - 1 : pat_class_of(E, N, d.odd));
    cls_out[p] = (unsigned char)d.cls;
    desc[p] = d;
    nchunks[p] = d.nchunks;
    if (d.cls == PAT_NCLS - 1) {
        atomicMax(&meta->maxE, E);
        atomicMax(&meta->maxN, N);
        if (d.odd >= 0) atomicMax(&meta->anyOdd, 1);
    }
}

// Exclusive scans of the per-pattern chunk counts, one per class (blockIdx.x = class c, one CTA of 1024 threads each):
// off[c][i] = chunks of the class-c patterns before pattern i, off[c][n] = totals[c] = all chunks of class c.
__global__ void __launch_bounds__(1024) scan_kernel(const unsigned int* __restrict__ in_all, const unsigned char* __restrict__ cls,
                                                    long long n, unsigned long long* __restrict__ off_all,
                                                    unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    const int c = blockIdx.x;
    unsigned long long* off = off_all + (size_t)c * (n + 1);
    const long long per = (n + 1023) / 1024;
    const long long lo = tid * per, hi = lo + per < n ? lo + per : n;
    unsigned long long s = 0;
    for (long long i = lo; i < hi; ++i) s += cls[i] == c ? in_all[i] : 0u;
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned long long v = tid >= d ? part[tid - d] : 0ull;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned long long run = tid ? part[tid - 1] : 0ull;
    for (long long i = lo; i < hi; ++i) { off[i] = run; run += cls[i] == c ? in_all[i] : 0u; }
    if (tid == 1023) { off[n] = part[1023]; totals[c] = part[1023]; }
}

struct PatParams {
    const double2* A;
    const double2* D;  // gamma(s) or null: n_gamma x nv
    const int32_t* gidx;  // per-pattern row of D, or null (all patterns use row 0)
    const int32_t* aidx;  // per-pattern matrix of the table A[n_A][nv][nv], or null (all patterns use matrix 0)
    int nv, glynn, smax, T, O;
    int lda, ldg;   // row strides of A and of the loop-vector table (>= nv: the samplers pass leading blocks of one matrix)
    long long B;
    const PatDesc* desc;
    const unsigned long long* coff;  // B + 1 chunk offsets
    unsigned long long nchunks;
    unsigned long long* counter;
    double* partial;  // nchunks * 4
};

}  // namespace wb
#include "pat_dmma.cuh"
namespace wb {

static_assert(PAT_NCLS == PD_NCLS + 1, "class table out of sync");
__host__ __device__ inline int pat_class_of(int E, int N, int odd) {
    // even and odd totals both run on the tensor-core kernel; the series order (N / 2, or N with an unpaired vertex) is bounded
    if ((odd >= 0 ? N : N / 2) > PD_TMAX) return PD_NCLS;
    return pd_class(E);
}

__global__ void __launch_bounds__(32 * BW_WARPS) pat_main_kernel(PatParams p) {
    extern __shared__ __align__(16) unsigned char smem_bw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wsb = warp_ws_bytes(p.smax, p.T, p.O);
    WarpWs w = carve(smem_bw + warp * wsb, p.smax, p.T, p.O);
    __shared__ int s_k[BW_WARPS], s_esum[BW_WARPS], s_d0[BW_WARPS];
    __shared__ double s_wt[BW_WARPS];
    const bool loops = p.D != nullptr;

    for (;;) {
        unsigned long long chunk = 0;
        if (lane == 0) chunk = atomicAdd(p.counter, 1ull);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= p.nchunks) break;
        // pattern owning this chunk: largest pat with coff[pat] <= chunk
        long long lo = 0, hi = p.B;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(p.coff + mid) <= chunk) lo = mid; else hi = mid;
        }
        const long long pat = lo;
        const PatDesc* d = p.desc + pat;
        const int E = d->E, N = d->N, odd = d->odd;
        const unsigned long long steps = d->steps;
        const double2* Dp = (p.D && p.gidx) ? p.D + (size_t)__ldg(p.gidx + pat) * p.ldg : p.D;
        const double2* Ap = p.aidx ? p.A + (size_t)__ldg(p.aidx + pat) * p.lda * p.lda : p.A;
        __syncwarp();
        if (lane < BW_EMAX) { w.eu[lane] = d->u[lane]; w.ev[lane] = d->v[lane]; w.er[lane] = d->r[lane]; }
        __syncwarp();
        const unsigned long long jb = (chunk - __ldg(p.coff + pat)) * BW_CHUNK;
        const unsigned long long je = jb + BW_CHUNK < steps ? jb + BW_CHUNK : steps;
        const int T = N / 2, order = odd >= 0 ? N : N / 2;
        cdd acc;
        acc.re = {0.0, 0.0};
        acc.im = {0.0, 0.0};
        for (unsigned long long j = jb; j < je; ++j) {
            if (lane == 0) {
                int es, d0;
                double wt;
                s_k[warp] = decode_subset(w, E, p.glynn, j, 0, &es, &wt, &d0);
                s_esum[warp] = es; s_wt[warp] = wt; s_d0[warp] = d0;
            }
            __syncwarp();
            const int k = s_k[warp];
            subset_traces(w, Ap, p.lda, Dp, odd, -1, k, T, lane);
            if (odd >= 0) fac_odd(w, order, __ldg(Dp + odd), w.ov, lane);
            else fac_even(w, order, lane);
            exp_series(w, w.cs0, order, lane);
            if (lane == 0) {
                double pre = (((N / 2 - s_esum[warp]) & 1) ? -1.0 : 1.0) * s_wt[warp];
                if (p.glynn && odd < 0 && s_d0[warp]) pre *= 0.5;
                dd_add(acc.re, pre * w.cs0[order].x);
                dd_add(acc.im, pre * w.cs0[order].y);
            }
            __syncwarp();
        }
        if (lane == 0) {
            double* o = p.partial + chunk * 4;
            o[0] = acc.re.hi; o[1] = acc.re.lo; o[2] = acc.im.hi; o[3] = acc.im.lo;
        }
        (void)loops;
    }
}

struct PatBases {
    unsigned long long base[PAT_NCLS];   // first slot of class c in the partial table
};

__global__ void pat_final_kernel(const PatDesc* __restrict__ desc, const unsigned long long* __restrict__ coff_all, PatBases cb,
                                 const double* __restrict__ partial_all, const double2* __restrict__ D,
                                 const int32_t* __restrict__ gidx, int ldg, int glynn, long long B,
                                 double2* __restrict__ out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const PatDesc* d = desc + p;
    if (d->kind == 1) { out[p] = make_double2(1.0, 0.0); return; }
    if (d->kind == 2) { out[p] = make_double2(0.0, 0.0); return; }
    if (d->kind == 3) { out[p] = D[(gidx ? (size_t)gidx[p] * ldg : 0) + d->odd]; return; }
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    const unsigned long long* coff = coff_all + (size_t)d->cls * (B + 1);
    const double* partial = partial_all + 4 * cb.base[d->cls];
    for (unsigned long long c = coff[p]; c < coff[p + 1]; ++c) {
        dd_add_dd(re, dd{partial[c * 4 + 0], partial[c * 4 + 1]});
        dd_add_dd(im, dd{partial[c * 4 + 2], partial[c * 4 + 3]});
    }
    double scale = 1.0;
    if (glynn) scale = ldexp(1.0, -(d->odd >= 0 ? d->N / 2 : d->N / 2 - 1));  // _hafnian.py:571-575
    out[p] = make_double2((re.hi + re.lo) * scale, (im.hi + im.lo) * scale);
}

// =================================================================================================
// loop_hafnian_batch sweep
// =================================================================================================
struct BatchParams {
    const double2* A;   // n x n, edge ordered (vertex e paired with e + E)
    const double2* D;   // n_D x n loop vectors (loop_hafnian_batch: n_D = 1; ..._gamma: one row per displacement)
    int n, n_D, E, glynn, odd_variant, N_fixed, N_max, length, smax, T, O;
    int reps[BW_EMAX];
    unsigned long long j0, j1;
    double* partials;   // (gridDim.x * warps per CTA) x n_D x length x 4, zero-initialised
};

// One warp per subset.  The reduced matrix and its power traces are built once per subset and shared by all
// n_D loop vectors (the reference recomputes only XD_S, D_S per vector too: loop_hafnian_batch_gamma.py:107-111);
// the warp's accumulators live in its own slice of global memory (L2 resident), double-double, no atomics.
__global__ void __launch_bounds__(32 * BW_WARPS) batch_kernel(BatchParams p) {
    extern __shared__ __align__(16) unsigned char smem_bw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wsb = warp_ws_bytes(p.smax, p.T, p.O);
    WarpWs w = carve(smem_bw + warp * wsb, p.smax, p.T, p.O);
    __shared__ int s_k[BW_WARPS], s_esum[BW_WARPS], s_d0[BW_WARPS];
    __shared__ double s_wt[BW_WARPS];
    const int E = p.E;
    if (lane < BW_EMAX) {
        w.eu[lane] = (unsigned char)lane; w.ev[lane] = (unsigned char)(lane + E);
        w.er[lane] = (unsigned short)(lane < E ? p.reps[lane] : 0);
    }
    __syncwarp();
    const int T = p.N_max / 2;
    const int wpc = blockDim.x >> 5;
    const unsigned long long gw = (unsigned long long)blockIdx.x * wpc + warp, nw = (unsigned long long)gridDim.x * wpc;
    double* acc_all = p.partials + gw * 4 * (size_t)p.length * p.n_D;
    for (unsigned long long j = p.j0 + gw; j < p.j1; j += nw) {
        if (lane == 0) {
            int es, d0;
            double wt;
            s_k[warp] = decode_subset(w, E, p.glynn, j, 1, &es, &wt, &d0);
            s_esum[warp] = es; s_wt[warp] = wt; s_d0[warp] = d0;
        }
        __syncwarp();
        const int k = s_k[warp], esum = s_esum[warp];
        const double wt = s_wt[warp];
        const int kept0 = w.kept[0];
        const bool extra = p.odd_variant && kept0 == 0 && w.kept[1] == 0;
        subset_setup(w, p.A, p.n, 0, extra ? 1 : -1, k, T, lane);
        for (int dk = 0; dk < p.n_D; ++dk) {
            const double2* Dk = p.D + (size_t)dk * p.n;
            double* acc = acc_all + (size_t)dk * 4 * p.length;      // [length][4]
            const double2 oddloop = __ldg(Dk + 0), oddloop0 = p.odd_variant ? __ldg(Dk + 1) : make_double2(0.0, 0.0);
            subset_loops(w, Dk, true, extra, k, T, lane);
            fac_even(w, p.N_max / 2, lane);
            exp_series(w, w.cs0, p.N_max / 2, lane);           // f_loop
            fac_odd(w, p.N_max, oddloop, w.ov, lane);
            exp_series(w, w.cs1, p.N_max, lane);               // f_loop_odd
            if (extra) {                                        // loop_hafnian_batch.py:181-185
                fac_odd(w, p.N_fixed, oddloop0, w.ov0, lane);
                exp_series(w, w.cs2, p.N_fixed, lane);
                if (lane == 0) {
                    const double pm = ((p.N_fixed / 2 - esum) & 1) ? -1.0 : 1.0;
                    dd a = {acc[0], acc[1]}, b = {acc[2], acc[3]};
                    dd_add(a, wt * pm * w.cs2[p.N_fixed].x);
                    dd_add(b, wt * pm * w.cs2[p.N_fixed].y);
                    acc[0] = a.hi; acc[1] = a.lo; acc[2] = b.hi; acc[3] = b.lo;
                }
                __syncwarp();
            }
            const int first = 2 * kept0 + (p.odd_variant ? 1 : 0);
            for (int nd = first + lane; nd < p.length; nd += 32) {   // loop_hafnian_batch.py:105-114, 188-200
                const int N = p.N_fixed + nd;
                const double pm = ((N / 2 - esum) & 1) ? -1.0 : 1.0;
                const int half = p.odd_variant ? (nd - 1) / 2 : nd / 2;
                const double wgt = binom_d(half, kept0) * wt * pm;
                const double2 v = (N & 1) ? w.cs1[N] : w.cs0[N / 2];
                dd a = {acc[nd * 4 + 0], acc[nd * 4 + 1]}, b = {acc[nd * 4 + 2], acc[nd * 4 + 3]};
                dd_add(a, wgt * v.x);
                dd_add(b, wgt * v.y);
                acc[nd * 4 + 0] = a.hi; acc[nd * 4 + 1] = a.lo; acc[nd * 4 + 2] = b.hi; acc[nd * 4 + 3] = b.lo;
            }
            __syncwarp();
        }
    }
}

// out[nd][4] = fixed-order sum over warps
__global__ void batch_final_kernel(const double* __restrict__ partials, int nwarps, int length, double* __restrict__ out) {
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= length) return;
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    for (int wv = 0; wv < nwarps; ++wv) {
        const double* q = partials + ((size_t)wv * length + nd) * 4;
        dd_add_dd(re, dd{q[0], q[1]});
        dd_add_dd(im, dd{q[2], q[3]});
    }
    out[nd * 4 + 0] = re.hi; out[nd * 4 + 1] = re.lo; out[nd * 4 + 2] = im.hi; out[nd * 4 + 3] = im.lo;
}

// =================================================================================================
// montrealer / loop montrealer (thewalrus/_montrealer.py:37-102)
// =================================================================================================
// mtl(A) = (-1)^(n+1) [ V / (2n) + W / 2 ],  V = sum_p (-1)^(|p|+1) tr(Sigma_p^n),
// W = sum_p (-1)^(|p|+1) conj(zeta_p) Sigma_p^(n-1) zeta_p,  Sigma = X A,  p over the non-empty mode subsets
// (label bit i, MSB first, selects mode i: dec2bin :17-34).  With R = p u (p + n) and P the pair swap on R,
// Sigma_p = P A_RR, so tr(Sigma_p^n) = tr((A_RR P)^n) — the matrix subset_setup builds with delta = 1 — and
// conj(zeta_p) Sigma_p^(n-1) zeta_p = (P conj(zeta_R)) (A_RR P)^(n-1) (P zeta_R).  One warp per subset.
struct MtlParams {
    const double2* A;      // 2n x 2n
    const double2* zeta;   // 2n or null
    const double2* zetac;  // conj(zeta) or null
    int n, smax, T;
    unsigned long long p0, p1;
    double* partials;      // (gridDim.x * warps) x 8: V (re_hi, re_lo, im_hi, im_lo), W (same)
};

__global__ void __launch_bounds__(32 * BW_WARPS) mtl_kernel(MtlParams p) {
    extern __shared__ __align__(16) unsigned char smem_bw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wsb = warp_ws_bytes(p.smax, p.T, 0);
    WarpWs w = carve(smem_bw + warp * wsb, p.smax, p.T, 0);
    __shared__ int s_k[BW_WARPS];
    const int n = p.n;
    const int wpc = blockDim.x >> 5;
    const unsigned long long gw = (unsigned long long)blockIdx.x * wpc + warp, nw = (unsigned long long)gridDim.x * wpc;
    cdd V, W;
    V.re = {0.0, 0.0}; V.im = {0.0, 0.0}; W.re = {0.0, 0.0}; W.im = {0.0, 0.0};
    for (unsigned long long lab = p.p0 + gw; lab < p.p1; lab += nw) {
        if (lab == 0) continue;                       // the empty subset contributes nothing
        if (lane == 0) {
            int k = 0;
            for (int i = 0; i < n; ++i)
                if ((lab >> (n - 1 - i)) & 1ull) w.rows[k++] = i;
            for (int a = 0; a < k; ++a) { w.rows[k + a] = w.rows[a] + n; w.delta[a] = 1.0; w.delta[k + a] = 1.0; }
            s_k[warp] = k;
        }
        __syncwarp();
        const int k = s_k[warp];
        subset_setup(w, p.A, 2 * n, -1, -1, k, n, lane);
        if (p.zeta) subset_loops_lr(w, p.zetac, p.zeta, true, false, false, k, n, lane);
        if (lane == 0) {
            const double sg = (k & 1) ? 1.0 : -1.0;    // (-1)^(|p| + 1)
            // tr(M^n): ptr[n] for n >= 1 (ptr[1] is the plain trace)
            dd_add(V.re, sg * w.ptr[n].x);
            dd_add(V.im, sg * w.ptr[n].y);
            if (p.zeta) { dd_add(W.re, sg * w.lv[n].x); dd_add(W.im, sg * w.lv[n].y); }
        }
        __syncwarp();
    }
    if (lane == 0) {
        double* o = p.partials + gw * 8;
        o[0] = V.re.hi; o[1] = V.re.lo; o[2] = V.im.hi; o[3] = V.im.lo;
        o[4] = W.re.hi; o[5] = W.re.lo; o[6] = W.im.hi; o[7] = W.im.lo;
    }
}

struct DevBufB {
    void* p = nullptr;
    ~DevBufB() { if (p) pool_free(p); }
};
struct EvPair {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~EvPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

// Warps per CTA (1..BW_WARPS) maximising resident warps per SM for a per-warp shared-memory footprint;
// returns 0 if even one warp does not fit.  *ctas = CTAs per SM at that choice.
static int bw_pick_warps(size_t per_warp, int* ctas) {
    const size_t budget = 227 * 1024, per_cta_overhead = 1024 + 256;
    int best_w = 0, best_total = 0, best_c = 0;
    for (int wv = 1; wv <= BW_WARPS; ++wv) {
        const size_t cta = per_warp * wv + per_cta_overhead;
        if (cta > budget) break;
        int c = (int)(budget / cta);
        if (c > 32 / wv) c = 32 / wv;   // keep <= 32 warps per SM (register budget of these kernels)
        if (c < 1) c = 1;
        if (c * wv >= best_total) { best_total = c * wv; best_w = wv; best_c = c; }
    }
    *ctas = best_c;
    return best_w;
}

// Device-side driver of the batched front end: everything between "inputs are in HBM" and "results are in HBM" on the
// caller's stream.  Scratch comes from the caching pool.  ONE host synchronisation in the middle (the class totals and
// shared-memory maxima decide the launch shapes) and one at the end (the scratch is released on return).
// env WB200_PAT_DFMA=1 forces every pattern onto the warp-per-subset DFMA kernel (A/B measurements, tests).
static int lhaf_matrices_device(const double2* dA, int n_A, const int32_t* dai, const double2* dD, int n_gamma, const int32_t* dgi,
                                int nv, const int32_t* drpt, int64_t B, int glynn, double2* dout, int sms, cudaStream_t st,
                                double* kernel_ms, int lda = 0, int ldg = 0) {
    if (lda < nv) lda = nv;
    if (ldg < nv) ldg = nv;
    const char* env_dfma = getenv("WB200_PAT_DFMA");      // read per call: the tests toggle it
    const int force_fallback = (env_dfma && atoi(env_dfma)) ? 1 : 0;
    StreamBuf ddesc, dnch, dcls, dcoff, dmeta, dpartial;
    WB_POOL(ddesc.alloc(sizeof(PatDesc) * (size_t)B, st));
    WB_POOL(dnch.alloc(sizeof(unsigned int) * (size_t)B, st));
    WB_POOL(dcls.alloc((size_t)B, st));
    WB_POOL(dcoff.alloc(sizeof(unsigned long long) * ((size_t)B + 1) * PAT_NCLS, st));
    // meta block: PatMeta | totals[PAT_NCLS] | counters[PAT_NCLS]
    const size_t meta_bytes = 64 + 2 * sizeof(unsigned long long) * PAT_NCLS;
    WB_POOL(dmeta.alloc(meta_bytes, st));
    WB_CUDA(cudaMemsetAsync(dmeta.p, 0, meta_bytes, st));
    PatMeta* d_meta = (PatMeta*)dmeta.p;
    unsigned long long* d_totals = (unsigned long long*)((char*)dmeta.p + 64);
    unsigned long long* d_counters = d_totals + PAT_NCLS;
    EvPair ev;
    if (kernel_ms) {
        WB_CUDA(cudaEventCreate(&ev.e0));
        WB_CUDA(cudaEventCreate(&ev.e1));
        WB_CUDA(cudaEventRecord(ev.e0, st));
    }
    pat_prep_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(drpt, B, nv, dD != nullptr, glynn, force_fallback, dai, n_A, dgi,
                                                                 n_gamma, (PatDesc*)ddesc.p, (unsigned int*)dnch.p,
                                                                 (unsigned char*)dcls.p, d_meta);
    scan_kernel<<<PAT_NCLS, 1024, 0, st>>>((const unsigned int*)dnch.p, (const unsigned char*)dcls.p, B,
                                           (unsigned long long*)dcoff.p, d_totals);
    WB_CUDA(cudaGetLastError());
    struct { PatMeta meta; char pad[64 - sizeof(PatMeta)]; unsigned long long totals[PAT_NCLS]; } h;
    WB_CUDA(cudaMemcpyAsync(&h, dmeta.p, 64 + sizeof(unsigned long long) * PAT_NCLS, cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    if (h.meta.err == 3) { set_error("lhaf_patterns: A_index / gamma_index entry outside the table"); return WB200_EINVAL; }
    if (h.meta.err) {
        set_error(h.meta.err == 1 ? "lhaf_patterns: repetition counts must be in [0, 65535] with total <= 32000"
                                  : "lhaf_patterns: a pattern exceeds the kernel limits (edges <= %d, series order <= %d, steps <= 1e12) or has an odd total without loops", BW_EMAX, BW_MAX_ORDER);
        return h.meta.err == 1 ? WB200_EINVAL : WB200_ENOSUP;
    }
    PatBases cb;
    unsigned long long nchunks = 0;
    for (int c = 0; c < PAT_NCLS; ++c) { cb.base[c] = nchunks; nchunks += h.totals[c]; }
    if (nchunks > 0) {
        WB_POOL(dpartial.alloc(sizeof(double) * 4 * (size_t)nchunks, st));
        // largest classes first: the small ones fill the tail of the big ones' last wave
        for (int c = PAT_NCLS - 1; c >= 0; --c) {
            if (!h.totals[c]) continue;
            PatParams p;
            p.A = dA; p.D = dD; p.gidx = dgi; p.aidx = dai; p.nv = nv; p.glynn = glynn; p.lda = lda; p.ldg = ldg;
            p.smax = 2 * h.meta.maxE; p.T = h.meta.maxN / 2; p.O = h.meta.anyOdd ? h.meta.maxN : h.meta.maxN / 2;
            p.B = B; p.desc = (const PatDesc*)ddesc.p; p.coff = (const unsigned long long*)dcoff.p + (size_t)c * (B + 1);
            p.nchunks = h.totals[c]; p.counter = d_counters + c;
            p.partial = (double*)dpartial.p + 4 * cb.base[c];
            if (c < PAT_NCLS - 1) {
                int rc = launch_pat_class(c, p, sms, st);
                if (rc) return rc;
                continue;
            }
            const size_t per_warp = warp_ws_bytes(p.smax, p.T, p.O);
            int ctas = 1;
            const int wpc = bw_pick_warps(per_warp, &ctas);
            if (wpc < 1) { set_error("lhaf_patterns: pattern too large for shared memory (%zu bytes per warp)", per_warp); return WB200_ENOSUP; }
            const size_t shm = per_warp * wpc;
            WB_CUDA(cudaFuncSetAttribute(pat_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
            int grid = sms * ctas;
            const unsigned long long want = (p.nchunks + wpc - 1) / wpc;
            if ((unsigned long long)grid > want) grid = (int)want;
            pat_main_kernel<<<grid, 32 * wpc, shm, st>>>(p);
            WB_CUDA(cudaGetLastError());
        }
    }
    pat_final_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>((const PatDesc*)ddesc.p, (const unsigned long long*)dcoff.p, cb,
                                                                  (const double*)dpartial.p, dD, dgi, ldg, glynn, B, dout);
    WB_CUDA(cudaGetLastError());
    if (kernel_ms) {
        WB_CUDA(cudaEventRecord(ev.e1, st));
        WB_CUDA(cudaEventSynchronize(ev.e1));
        float ms = 0;
        WB_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        *kernel_ms = ms;
    }
    return WB200_OK;       // the stream-ordered scratch is released when the stream reaches this point
}

}  // namespace wb

using namespace wb;

extern "C" int wb200_lhaf_patterns_host(int device, const double* A, const double* gamma, int nv, const int32_t* rpt,
                                        int64_t B, int glynn, double* out, double* kernel_ms) {
    return wb200_lhaf_patterns_multi_host(device, A, gamma, gamma ? 1 : 0, nullptr, nv, rpt, B, glynn, out, kernel_ms);
}

extern "C" int wb200_lhaf_patterns_multi_host(int device, const double* A, const double* gamma, int n_gamma,
                                              const int32_t* gamma_index, int nv, const int32_t* rpt, int64_t B,
                                              int glynn, double* out, double* kernel_ms) {
    return wb200_lhaf_matrices_host(device, A, 1, nullptr, gamma, n_gamma, gamma_index, nv, rpt, B, glynn, out, kernel_ms);
}

static int pat_check_args(const void* A, int n_A, const void* A_index, const void* gamma, int n_gamma, const void* gamma_index,
                          int nv, const void* rpt, int64_t B, const void* out) {
    if (!A || !rpt || !out) { set_error("lhaf_patterns: null pointer"); return WB200_EINVAL; }
    if (n_A < 1 || (n_A > 1 && !A_index)) { set_error("lhaf_patterns: n_A must be >= 1 and A_index given for n_A > 1"); return WB200_EINVAL; }
    if (gamma && n_gamma < 1) { set_error("lhaf_patterns: n_gamma must be >= 1 when gamma is given"); return WB200_EINVAL; }
    if (gamma && n_gamma > 1 && !gamma_index) { set_error("lhaf_patterns: gamma_index is required for n_gamma > 1"); return WB200_EINVAL; }
    if (nv < 1 || nv > BW_NVMAX) { set_error("lhaf_patterns: %d vertices outside [1, %d]", nv, BW_NVMAX); return nv > BW_NVMAX ? WB200_ENOSUP : WB200_EINVAL; }
    if (B < 0) { set_error("lhaf_patterns: negative batch"); return WB200_EINVAL; }
    return WB200_OK;
}

extern "C" int wb200_lhaf_matrices_dev(const double* dA, int n_A, const int32_t* dA_index, const double* dgamma, int n_gamma,
                                       const int32_t* dgamma_index, int nv, const int32_t* drpt, int64_t B, int glynn,
                                       double* d_out, void* stream) {
    int rc = pat_check_args(dA, n_A, dA_index, dgamma, n_gamma, dgamma_index, nv, drpt, B, d_out);
    if (rc || B == 0) return rc;
    int device = 0, sms = 0;
    (void)cudaGetLastError();
    WB_CUDA(cudaGetDevice(&device));
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    return lhaf_matrices_device((const double2*)dA, n_A, dA_index, (const double2*)dgamma, n_gamma, dgamma ? dgamma_index : nullptr, nv,
                                drpt, B, glynn, (double2*)d_out, sms, (cudaStream_t)stream, nullptr);
}

extern "C" int wb200_lhaf_matrices_host(int device, const double* A, int n_A, const int32_t* A_index,
                                        const double* gamma, int n_gamma, const int32_t* gamma_index, int nv,
                                        const int32_t* rpt, int64_t B, int glynn, double* out, double* kernel_ms) {
    int rc = pat_check_args(A, n_A, A_index, gamma, n_gamma, gamma_index, nv, rpt, B, out);
    if (rc) return rc;
    if (A_index)      // host tables are checked before any CUDA call (the device twin checks them in its prep kernel)
        for (int64_t i = 0; i < B; ++i)
            if (A_index[i] < 0 || A_index[i] >= n_A) { set_error("lhaf_patterns: A_index[%lld] = %d outside [0, %d)", (long long)i, A_index[i], n_A); return WB200_EINVAL; }
    if (gamma && gamma_index)
        for (int64_t i = 0; i < B; ++i)
            if (gamma_index[i] < 0 || gamma_index[i] >= n_gamma) { set_error("lhaf_patterns: gamma_index[%lld] = %d outside [0, %d)", (long long)i, gamma_index[i], n_gamma); return WB200_EINVAL; }
    if (B == 0) { if (kernel_ms) *kernel_ms = 0.0; return WB200_OK; }
    WB_CUDA(cudaSetDevice(device));
    int sms = 0;
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    DevBufB dA, dai, dD, dgi, drpt, dout;
    WB_POOL(pool_alloc(&dA.p, sizeof(double2) * (size_t)n_A * nv * nv));
    WB_CUDA(cudaMemcpy(dA.p, A, sizeof(double2) * (size_t)n_A * nv * nv, cudaMemcpyHostToDevice));
    if (A_index) {
        WB_POOL(pool_alloc(&dai.p, sizeof(int32_t) * (size_t)B));
        WB_CUDA(cudaMemcpy(dai.p, A_index, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice));
    }
    if (gamma) {
        WB_POOL(pool_alloc(&dD.p, sizeof(double2) * (size_t)n_gamma * nv));
        WB_CUDA(cudaMemcpy(dD.p, gamma, sizeof(double2) * (size_t)n_gamma * nv, cudaMemcpyHostToDevice));
        if (gamma_index) {
            WB_POOL(pool_alloc(&dgi.p, sizeof(int32_t) * (size_t)B));
            WB_CUDA(cudaMemcpy(dgi.p, gamma_index, sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice));
        }
    }
    WB_POOL(pool_alloc(&drpt.p, sizeof(int32_t) * (size_t)B * nv));
    WB_CUDA(cudaMemcpy(drpt.p, rpt, sizeof(int32_t) * (size_t)B * nv, cudaMemcpyHostToDevice));
    WB_POOL(pool_alloc(&dout.p, sizeof(double2) * (size_t)B));
    rc = lhaf_matrices_device((const double2*)dA.p, n_A, (const int32_t*)dai.p, (const double2*)dD.p, n_gamma, (const int32_t*)dgi.p, nv,
                              (const int32_t*)drpt.p, B, glynn, (double2*)dout.p, sms, (cudaStream_t)0, kernel_ms);
    if (rc) return rc;
    WB_CUDA(cudaMemcpy(out, dout.p, sizeof(double2) * (size_t)B, cudaMemcpyDeviceToHost));
    return WB200_OK;
}

extern "C" int wb200_lhaf_batch_steps(const int32_t* edge_reps, int n_edges, uint64_t* steps) {
    if (!edge_reps || !steps || n_edges < 1) { set_error("lhaf_batch_steps: bad arguments"); return WB200_EINVAL; }
    unsigned __int128 s = 1;
    for (int i = 0; i < n_edges; ++i) {
        if (edge_reps[i] < 0) { set_error("negative edge repetition"); return WB200_EINVAL; }
        s *= (uint64_t)edge_reps[i] + 1;
        if (s > (((unsigned __int128)1) << 62)) { set_error("subset index space exceeds 2^62"); return WB200_ENOSUP; }
    }
    *steps = (uint64_t)s;
    return WB200_OK;
}

extern "C" int wb200_lhaf_batch_gamma_dev(const double* dAx, const double* dDx, int n, int n_D, const int32_t* edge_reps,
                                          int odd_variant, int cutoff_extra, int glynn, uint64_t j0, uint64_t j1, double* d_out,
                                          int length, void* stream) {
    if (!dAx || !dDx || !edge_reps || !d_out) { set_error("lhaf_batch: null pointer"); return WB200_EINVAL; }
    if (n < 2 || (n & 1)) { set_error("lhaf_batch: n must be even and >= 2 (got %d)", n); return WB200_EINVAL; }
    if (n_D < 1 || n_D > 65536) { set_error("lhaf_batch: number of loop vectors %d outside [1, 65536]", n_D); return WB200_EINVAL; }
    const int E = n / 2;
    if (E > BW_EMAX) { set_error("lhaf_batch: %d edges exceed the limit of %d", E, BW_EMAX); return WB200_ENOSUP; }
    if (odd_variant && (E < 2 || edge_reps[1] != 1)) { set_error("lhaf_batch: odd variant needs edge_reps[1] == 1"); return WB200_EINVAL; }
    uint64_t steps = 0;
    int rc = wb200_lhaf_batch_steps(edge_reps, E, &steps);
    if (rc) return rc;
    if (j0 > j1 || j1 > steps) { set_error("lhaf_batch: bad subset range"); return WB200_EINVAL; }
    BatchParams p;
    memset(&p, 0, sizeof(p));
    const int batch_max = edge_reps[0];
    int fixed_sum = 0;
    for (int i = 0; i < E; ++i) { p.reps[i] = edge_reps[i]; if (i >= (odd_variant ? 2 : 1)) fixed_sum += edge_reps[i]; }
    if (odd_variant) {
        p.N_fixed = 2 * fixed_sum + 1;
        p.N_max = p.N_fixed + 2 * batch_max + cutoff_extra + 1;
        p.length = 2 * batch_max + cutoff_extra + 2;
    } else {
        p.N_fixed = 2 * fixed_sum;
        p.N_max = p.N_fixed + 2 * batch_max + cutoff_extra;
        p.length = 2 * batch_max + cutoff_extra + 1;
    }
    if (length != p.length) { set_error("lhaf_batch: output length must be %d (got %d)", p.length, length); return WB200_EINVAL; }
    if (p.N_max > BW_MAX_ORDER) { set_error("lhaf_batch: photon number %d too large", p.N_max); return WB200_ENOSUP; }
    p.n = n; p.n_D = n_D; p.E = E; p.glynn = glynn; p.odd_variant = odd_variant; p.j0 = j0; p.j1 = j1;
    p.smax = n; p.T = p.N_max / 2; p.O = p.N_max;
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0, sms = 0;
    (void)cudaGetLastError();
    WB_CUDA(cudaGetDevice(&device));
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    p.A = (const double2*)dAx; p.D = (const double2*)dDx;
    const size_t per_warp = warp_ws_bytes(p.smax, p.T, p.O);
    int ctas = 1;
    const int wpc = bw_pick_warps(per_warp, &ctas);
    if (wpc < 1) { set_error("lhaf_batch: problem too large for shared memory (%zu bytes per warp)", per_warp); return WB200_ENOSUP; }
    const size_t shm = per_warp * wpc;
    WB_CUDA(cudaFuncSetAttribute(batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    int grid = sms * ctas;
    const uint64_t total = j1 - j0, want = (total + wpc - 1) / wpc;
    if ((uint64_t)grid > want) grid = (int)(want ? want : 1);
    const size_t rows = (size_t)p.length * n_D;                 // independent outputs
    while (grid > 1 && rows * (size_t)grid * wpc * 32 > ((size_t)1 << 31)) grid /= 2;   // bound the partial table (2 GiB)
    const int nwarps = grid * wpc;
    StreamBuf dpart;
    WB_POOL(dpart.alloc(sizeof(double) * 4 * rows * nwarps, st));
    WB_CUDA(cudaMemsetAsync(dpart.p, 0, sizeof(double) * 4 * rows * nwarps, st));
    p.partials = (double*)dpart.p;
    batch_kernel<<<grid, 32 * wpc, shm, st>>>(p);
    batch_final_kernel<<<(unsigned)((rows + 63) / 64), 64, 0, st>>>((const double*)dpart.p, nwarps, (int)rows, d_out);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

namespace {
struct HostTimer {       // CUDA events around a *_dev call on the legacy stream, only when the caller asks for the time
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~HostTimer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    int start(bool want) {
        if (!want) return WB200_OK;
        WB_CUDA(cudaEventCreate(&e0)); WB_CUDA(cudaEventCreate(&e1)); WB_CUDA(cudaEventRecord(e0, 0));
        return WB200_OK;
    }
    int stop(double* ms) {
        if (!e0) return WB200_OK;
        WB_CUDA(cudaEventRecord(e1, 0));
        WB_CUDA(cudaEventSynchronize(e1));
        float f = 0;
        WB_CUDA(cudaEventElapsedTime(&f, e0, e1));
        if (ms) *ms = f;
        return WB200_OK;
    }
};
}  // namespace

extern "C" int wb200_lhaf_batch_gamma_host(int device, const double* Ax, const double* Dx, int n, int n_D,
                                           const int32_t* edge_reps, int odd_variant, int cutoff_extra, int glynn,
                                           uint64_t j0, uint64_t j1, double* out, int length, double* kernel_ms) {
    if (!Ax || !Dx || !edge_reps || !out) { set_error("lhaf_batch: null pointer"); return WB200_EINVAL; }
    if (n < 2 || (n & 1) || n > 2 * BW_EMAX || n_D < 1 || n_D > 65536 || length < 1) {
        set_error("lhaf_batch: bad sizes (n = %d, n_D = %d, length = %d)", n, n_D, length);
        return n > 2 * BW_EMAX ? WB200_ENOSUP : WB200_EINVAL;
    }
    WB_CUDA(cudaSetDevice(device));
    DevBufB dA, dD, dout;
    const size_t rows = (size_t)length * n_D;
    WB_POOL(pool_alloc(&dA.p, sizeof(double2) * n * n));
    WB_CUDA(cudaMemcpy(dA.p, Ax, sizeof(double2) * n * n, cudaMemcpyHostToDevice));
    WB_POOL(pool_alloc(&dD.p, sizeof(double2) * (size_t)n * n_D));
    WB_CUDA(cudaMemcpy(dD.p, Dx, sizeof(double2) * (size_t)n * n_D, cudaMemcpyHostToDevice));
    WB_POOL(pool_alloc(&dout.p, sizeof(double) * 4 * rows));
    HostTimer tm;
    int rc = tm.start(kernel_ms != nullptr);
    if (rc) return rc;
    rc = wb200_lhaf_batch_gamma_dev((const double*)dA.p, (const double*)dD.p, n, n_D, edge_reps, odd_variant, cutoff_extra, glynn,
                                    j0, j1, (double*)dout.p, length, nullptr);
    if (rc) return rc;
    if ((rc = tm.stop(kernel_ms))) return rc;
    WB_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * 4 * rows, cudaMemcpyDeviceToHost));
    return WB200_OK;
}

extern "C" int wb200_lhaf_batch_host(int device, const double* Ax, const double* Dx, int n, const int32_t* edge_reps,
                                     int odd_variant, int cutoff_extra, int glynn, uint64_t j0, uint64_t j1, double* out,
                                     int length, double* kernel_ms) {
    return wb200_lhaf_batch_gamma_host(device, Ax, Dx, n, 1, edge_reps, odd_variant, cutoff_extra, glynn, j0, j1, out,
                                       length, kernel_ms);
}

__global__ void conj_kernel(const double2* __restrict__ in, int n, double2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(in[i].x, -in[i].y);
}

extern "C" int wb200_mtl_dev(const double* dA, const double* dzeta, int n_modes, uint64_t p0, uint64_t p1, double* d_out8,
                             void* stream) {
    if (!dA || !d_out8) { set_error("mtl: null pointer"); return WB200_EINVAL; }
    if (n_modes < 1 || 2 * n_modes > BW_NVMAX) {
        set_error("mtl: %d modes outside [1, %d]", n_modes, BW_NVMAX / 2);
        return n_modes > BW_NVMAX / 2 ? WB200_ENOSUP : WB200_EINVAL;
    }
    const int n = n_modes, n2 = 2 * n;
    const uint64_t total = 1ull << n;
    if (p0 > p1 || p1 > total) { set_error("mtl: bad subset range"); return WB200_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0, sms = 0;
    (void)cudaGetLastError();
    WB_CUDA(cudaGetDevice(&device));
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    MtlParams p;
    memset(&p, 0, sizeof(p));
    p.A = (const double2*)dA; p.n = n; p.smax = n2; p.T = n; p.p0 = p0; p.p1 = p1;
    StreamBuf dzc, dpart;
    if (dzeta) {
        WB_POOL(dzc.alloc(sizeof(double2) * n2, st));
        conj_kernel<<<1, 128, 0, st>>>((const double2*)dzeta, n2, (double2*)dzc.p);
        p.zeta = (const double2*)dzeta; p.zetac = (const double2*)dzc.p;
    }
    const size_t per_warp = warp_ws_bytes(p.smax, p.T, 0);
    int ctas = 1;
    const int wpc = bw_pick_warps(per_warp, &ctas);
    if (wpc < 1) { set_error("mtl: problem too large for shared memory (%zu bytes per warp)", per_warp); return WB200_ENOSUP; }
    const size_t shm = per_warp * wpc;
    WB_CUDA(cudaFuncSetAttribute(mtl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    int grid = sms * ctas;
    const uint64_t want = (p1 - p0 + wpc - 1) / wpc;
    if ((uint64_t)grid > want) grid = (int)(want ? want : 1);
    const int nwarps = grid * wpc;
    WB_POOL(dpart.alloc(sizeof(double) * 8 * nwarps, st));
    p.partials = (double*)dpart.p;
    mtl_kernel<<<grid, 32 * wpc, shm, st>>>(p);
    batch_final_kernel<<<1, 64, 0, st>>>((const double*)dpart.p, nwarps, 2, d_out8);   // two complex outputs
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

extern "C" int wb200_mtl_host(int device, const double* A, const double* zeta, int n_modes, uint64_t p0, uint64_t p1,
                              double out8[8], double* kernel_ms) {
    if (!A || !out8) { set_error("mtl: null pointer"); return WB200_EINVAL; }
    if (n_modes < 1 || 2 * n_modes > BW_NVMAX) {
        set_error("mtl: %d modes outside [1, %d]", n_modes, BW_NVMAX / 2);
        return n_modes > BW_NVMAX / 2 ? WB200_ENOSUP : WB200_EINVAL;
    }
    const int n2 = 2 * n_modes;
    WB_CUDA(cudaSetDevice(device));
    DevBufB dA, dz, dout;
    WB_POOL(pool_alloc(&dA.p, sizeof(double2) * n2 * n2));
    WB_CUDA(cudaMemcpy(dA.p, A, sizeof(double2) * n2 * n2, cudaMemcpyHostToDevice));
    if (zeta) {
        WB_POOL(pool_alloc(&dz.p, sizeof(double2) * n2));
        WB_CUDA(cudaMemcpy(dz.p, zeta, sizeof(double2) * n2, cudaMemcpyHostToDevice));
    }
    WB_POOL(pool_alloc(&dout.p, sizeof(double) * 8));
    HostTimer tm;
    int rc = tm.start(kernel_ms != nullptr);
    if (rc) return rc;
    rc = wb200_mtl_dev((const double*)dA.p, (const double*)dz.p, n_modes, p0, p1, (double*)dout.p, nullptr);
    if (rc) return rc;
    if ((rc = tm.stop(kernel_ms))) return rc;
    WB_CUDA(cudaMemcpy(out8, dout.p, sizeof(double) * 8, cudaMemcpyDeviceToHost));
    return WB200_OK;
}

// =================================================================================================
// chain-rule photon-number sampler, every mode step on the device
// =================================================================================================
// Replaces the per-mode loop of generate_hafnian_sample (thewalrus/samples.py:204-261) for S chains at once: at mode i a
// chain with outcomes n_1 .. n_(i-1) draws n_i from  p(k) ~ |lhaf(B[:i,:i], gamma[:i], reps = (n_1 .. n_(i-1), k))|^2 / k!,
// k = 0 .. cutoff, after the heterodyne shift  gamma <- gamma - het_i B[:, i]  (:243-249).  The S (cutoff + 1) loop
// hafnians of a mode step are one call of the batched front end on leading blocks of B and gamma (row strides M), the
// outcome is drawn on the device from a host-supplied uniform (one per chain and mode, numpy.random's stream), and
// only the final patterns travel back.
namespace wb {

__global__ void chain_update_kernel(double2* __restrict__ gamma, const double2* __restrict__ het, const double2* __restrict__ Bm,
                                    int mode, long long S, int M) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * M) return;
    const long long s = idx / M;
    const int j = (int)(idx - s * M);
    const double2 h = het[s * M + mode], b = Bm[(size_t)j * M + mode];
    double2 g = gamma[idx];
    g.x -= h.x * b.x - h.y * b.y;
    g.y -= h.x * b.y + h.y * b.x;
    gamma[idx] = g;
}

// rpt[(s K + k), :] = (det[s, 0 .. mode-1], k);  gidx[s K + k] = s
__global__ void chain_rpt_kernel(const int32_t* __restrict__ det, int32_t* __restrict__ rpt, int32_t* __restrict__ gidx,
                                 int mode, int K, long long S, int M) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= S * K) return;
    const long long s = row / K;
    const int k = (int)(row - s * K), m = mode + 1;
    for (int j = 0; j < mode; ++j) rpt[row * m + j] = det[s * M + j];
    rpt[row * m + mode] = k;
    gidx[row] = (int32_t)s;
}

struct ChainFact { double inv[64]; };   // 1 / k!

// det[s, mode] = inverse-CDF draw from p(k) = |lh[s, k]|^2 / k! with the uniform u[s]: the outcome
// numpy.random.choice(K, p = p / sum p) returns for that uniform (count of cdf entries <= u, samples.py:245-249).
__global__ void chain_draw_kernel(const double2* __restrict__ lh, const double* __restrict__ u, ChainFact f, int32_t* __restrict__ det,
                                  int mode, int K, long long S, int M, int* __restrict__ err) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double tot = 0.0;
    bool bad = false;
    for (int k = 0; k < K; ++k) {
        const double2 v = lh[s * K + k];
        const double p = (v.x * v.x + v.y * v.y) * f.inv[k];
        if (!(p >= 0.0) || isinf(p)) bad = true;
        tot += p;
    }
    if (bad || !(tot > 0.0)) { atomicExch(err, 1); det[s * M + mode] = 0; return; }
    double run = 0.0, cdf_last = 0.0;
    for (int k = 0; k < K; ++k) { const double2 v = lh[s * K + k]; cdf_last += ((v.x * v.x + v.y * v.y) * f.inv[k]) / tot; }
    int cnt = 0;
    const double us = u[s];
    for (int k = 0; k < K; ++k) {
        const double2 v = lh[s * K + k];
        run += ((v.x * v.x + v.y * v.y) * f.inv[k]) / tot;
        if (run / cdf_last <= us) ++cnt;
    }
    det[s * M + mode] = cnt < K - 1 ? cnt : K - 1;
}

}  // namespace wb

extern "C" int wb200_hafnian_chains_host(int device, const double* Bmat, const double* gamma0, const double* het,
                                         const double* uniforms, int M, int64_t S, int cutoff, int32_t* det_out,
                                         double* kernel_ms) {
    if (!Bmat || !gamma0 || !het || !uniforms || !det_out) { set_error("hafnian_chains: null pointer"); return WB200_EINVAL; }
    if (M < 1 || M > BW_NVMAX) { set_error("hafnian_chains: %d modes outside [1, %d]", M, BW_NVMAX); return M > BW_NVMAX ? WB200_ENOSUP : WB200_EINVAL; }
    if (S < 1 || cutoff < 0 || cutoff > 63) { set_error("hafnian_chains: need S >= 1 and 0 <= cutoff <= 63"); return WB200_EINVAL; }
    const int K = cutoff + 1;
    WB_CUDA(cudaSetDevice(device));
    int sms = 0;
    if (device_sm_count(device, &sms)) return WB200_ECUDA;
    cudaStream_t st = 0;
    DevBufB dB, dg, dh, du, ddet, drpt, dgi, dlh, derr;
    WB_POOL(pool_alloc(&dB.p, sizeof(double2) * (size_t)M * M));
    WB_POOL(pool_alloc(&dg.p, sizeof(double2) * (size_t)S * M));
    WB_POOL(pool_alloc(&dh.p, sizeof(double2) * (size_t)S * M));
    WB_POOL(pool_alloc(&du.p, sizeof(double) * (size_t)S * M));
    WB_POOL(pool_alloc(&ddet.p, sizeof(int32_t) * (size_t)S * M));
    WB_POOL(pool_alloc(&drpt.p, sizeof(int32_t) * (size_t)S * K * M));
    WB_POOL(pool_alloc(&dgi.p, sizeof(int32_t) * (size_t)S * K));
    WB_POOL(pool_alloc(&dlh.p, sizeof(double2) * (size_t)S * K));
    WB_POOL(pool_alloc(&derr.p, sizeof(int)));
    WB_CUDA(cudaMemcpy(dB.p, Bmat, sizeof(double2) * (size_t)M * M, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(dg.p, gamma0, sizeof(double2) * (size_t)S * M, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(dh.p, het, sizeof(double2) * (size_t)S * M, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemcpy(du.p, uniforms, sizeof(double) * (size_t)S * M, cudaMemcpyHostToDevice));
    WB_CUDA(cudaMemsetAsync(ddet.p, 0, sizeof(int32_t) * (size_t)S * M, st));
    WB_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
    ChainFact f;
    f.inv[0] = 1.0;
    for (int k = 1; k < 64; ++k) f.inv[k] = f.inv[k - 1] / k;
    HostTimer tm;
    int rc = tm.start(kernel_ms != nullptr);
    if (rc) return rc;
    const long long SK = (long long)S * K;
    for (int mode = 0; mode < M; ++mode) {
        chain_update_kernel<<<(unsigned)((S * M + 255) / 256), 256, 0, st>>>((double2*)dg.p, (const double2*)dh.p, (const double2*)dB.p,
                                                                            mode, S, M);
        chain_rpt_kernel<<<(unsigned)((SK + 127) / 128), 128, 0, st>>>((const int32_t*)ddet.p, (int32_t*)drpt.p, (int32_t*)dgi.p, mode, K,
                                                                     S, M);
        WB_CUDA(cudaGetLastError());
        rc = lhaf_matrices_device((const double2*)dB.p, 1, nullptr, (const double2*)dg.p, (int)(S < 0x7fffffff ? S : 0x7fffffff),
                                  (const int32_t*)dgi.p, mode + 1, (const int32_t*)drpt.p, SK, 1, (double2*)dlh.p, sms, st, nullptr, M, M);
        if (rc) return rc;
        chain_draw_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>((const double2*)dlh.p, (const double*)du.p + (size_t)mode * S, f,
                                                                    (int32_t*)ddet.p, mode, K, S, M, (int*)derr.p);
        WB_CUDA(cudaGetLastError());
    }
    if ((rc = tm.stop(kernel_ms))) return rc;
    int herr = 0;
    WB_CUDA(cudaMemcpy(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
    WB_CUDA(cudaMemcpy(det_out, ddet.p, sizeof(int32_t) * (size_t)S * M, cudaMemcpyDeviceToHost));
    if (herr) { set_error("hafnian_chains: probabilities contain NaN, are negative or do not sum to a positive number"); return WB200_EINVAL; }
    return WB200_OK;
}
