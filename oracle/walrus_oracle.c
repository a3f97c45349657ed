/* CPU ORACLE (C restatement) — TEST / BASELINE INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * Plain-C, OpenMP restatement of the reference's hot path (XanaduAI/thewalrus v0.23.0-dev) for the cases
 * the headline configurations use: all edge repetitions 1, Glynn sieve.  It follows the reference's
 * ALGORITHM (full product chain for the power traces, the `comb` exponential-series loop, Gray-code
 * permanent, Cholesky-updating torontonian recursion), not the GPU's:
 *   oracle_hafnian_range       thewalrus/_hafnian.py:416-467 + charpoly.py:301-318 + _hafnian.py:183-209
 *   oracle_loop_hafnian_range  thewalrus/_hafnian.py:512-577 + :212-242
 *   oracle_perm_range          thewalrus/_permanent.py:86-168
 *   oracle_tor_recursive       thewalrus/_torontonian.py:157-247
 *   oracle_tor_direct_range    thewalrus/_torontonian.py:123-154
 *   oracle_ltor_direct_range   thewalrus/_torontonian.py:369-412 (numba_ltor; parallel in the reference too)
 *   oracle_brs_range           thewalrus/_permanent.py:198-249 (brs / ubrs; the reference loop is serial)
 *   oracle_lhaf_patterns       thewalrus/_hafnian.py:80-159, 162-180, 212-285, 315-356, 512-631: loop hafnians with
 *                              repeated vertices (the per-pattern call of quantum/fock_tensors.py:191-232), one
 *                              pattern per OpenMP task; power traces by the product chain for every order (the
 *                              reference switches to La Budde beyond the matrix size, charpoly.py:319-326: same
 *                              numbers)
 * Built twice by oracle/build.py: REAL = double (liboracle.so: checker + CPU baseline) and REAL = long
 * double (liboracle_ld.so: extended-precision yardstick for sampled ranges at n = 50/56).
 * Parity status: PINNED — tests/test_oracle_golden.py checks these against the committed outputs of the
 * reference itself (tests/golden/reference_outputs.json) and its known-answer tests.
 * Used only by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_LONG_DOUBLE
typedef long double REAL;
#define SQRT sqrtl
#else
typedef double REAL;
#define SQRT sqrt
#endif

#define NPAD(n) (((n) + 7) & ~7)

/* C = A * B for complex matrices stored as separate re/im planes with leading dimension ld (multiple of 8).
 * Register-blocked (4 rows x 16 columns) so the CPU baseline is not handicapped against the reference's
 * BLAS zgemm (`A @ H`, thewalrus/charpoly.py:317). */
#ifndef ORACLE_LONG_DOUBLE
typedef double v8d __attribute__((vector_size(64), aligned(8)));
static void cmatmul(int n, int ld, const REAL* restrict Ar, const REAL* restrict Ai, const REAL* restrict Br,
                    const REAL* restrict Bi, REAL* restrict Cr, REAL* restrict Ci) {
    for (int i0 = 0; i0 < n; i0 += 4) {
        const int ib = (n - i0) < 4 ? (n - i0) : 4;
        for (int j0 = 0; j0 < ld; j0 += 8) {
            v8d cr0 = {0}, cr1 = {0}, cr2 = {0}, cr3 = {0}, ci0 = {0}, ci1 = {0}, ci2 = {0}, ci3 = {0};
            const REAL* a0r = Ar + (size_t)(i0 + 0) * ld; const REAL* a0i = Ai + (size_t)(i0 + 0) * ld;
            const REAL* a1r = Ar + (size_t)(i0 + (ib > 1 ? 1 : 0)) * ld; const REAL* a1i = Ai + (size_t)(i0 + (ib > 1 ? 1 : 0)) * ld;
            const REAL* a2r = Ar + (size_t)(i0 + (ib > 2 ? 2 : 0)) * ld; const REAL* a2i = Ai + (size_t)(i0 + (ib > 2 ? 2 : 0)) * ld;
            const REAL* a3r = Ar + (size_t)(i0 + (ib > 3 ? 3 : 0)) * ld; const REAL* a3i = Ai + (size_t)(i0 + (ib > 3 ? 3 : 0)) * ld;
            for (int k = 0; k < n; ++k) {
                const v8d br = *(const v8d*)(Br + (size_t)k * ld + j0);
                const v8d bi = *(const v8d*)(Bi + (size_t)k * ld + j0);
                cr0 += a0r[k] * br - a0i[k] * bi; ci0 += a0r[k] * bi + a0i[k] * br;
                cr1 += a1r[k] * br - a1i[k] * bi; ci1 += a1r[k] * bi + a1i[k] * br;
                cr2 += a2r[k] * br - a2i[k] * bi; ci2 += a2r[k] * bi + a2i[k] * br;
                cr3 += a3r[k] * br - a3i[k] * bi; ci3 += a3r[k] * bi + a3i[k] * br;
            }
            *(v8d*)(Cr + (size_t)(i0 + 0) * ld + j0) = cr0; *(v8d*)(Ci + (size_t)(i0 + 0) * ld + j0) = ci0;
            if (ib > 1) { *(v8d*)(Cr + (size_t)(i0 + 1) * ld + j0) = cr1; *(v8d*)(Ci + (size_t)(i0 + 1) * ld + j0) = ci1; }
            if (ib > 2) { *(v8d*)(Cr + (size_t)(i0 + 2) * ld + j0) = cr2; *(v8d*)(Ci + (size_t)(i0 + 2) * ld + j0) = ci2; }
            if (ib > 3) { *(v8d*)(Cr + (size_t)(i0 + 3) * ld + j0) = cr3; *(v8d*)(Ci + (size_t)(i0 + 3) * ld + j0) = ci3; }
        }
    }
}
#else
static void cmatmul(int n, int ld, const REAL* restrict Ar, const REAL* restrict Ai, const REAL* restrict Br,
                    const REAL* restrict Bi, REAL* restrict Cr, REAL* restrict Ci) {
    for (int i = 0; i < n; ++i) {
        REAL* restrict cr = Cr + (size_t)i * ld;
        REAL* restrict ci = Ci + (size_t)i * ld;
        for (int j = 0; j < ld; ++j) { cr[j] = 0; ci[j] = 0; }
        for (int k = 0; k < n; ++k) {
            const REAL ar = Ar[(size_t)i * ld + k], ai = Ai[(size_t)i * ld + k];
            const REAL* restrict br = Br + (size_t)k * ld;
            const REAL* restrict bi = Bi + (size_t)k * ld;
            for (int j = 0; j < ld; ++j) {
                cr[j] += ar * br[j] - ai * bi[j];
                ci[j] += ar * bi[j] + ai * br[j];
            }
        }
    }
}
#endif

/* coefficients 0..order of exp(sum_i fac[i] eta^i): the reference's comb loop (_hafnian.py:196-209) */
static void exp_series(int order, const REAL* facr, const REAL* faci, REAL* outr, REAL* outi, REAL* tmpr, REAL* tmpi) {
    REAL* cr = outr; REAL* ci = outi; REAL* nr = tmpr; REAL* ni = tmpi;
    for (int k = 0; k <= order; ++k) { cr[k] = 0; ci[k] = 0; }
    cr[0] = 1;
    for (int i = 1; i <= order; ++i) {
        memcpy(nr, cr, sizeof(REAL) * (order + 1));
        memcpy(ni, ci, sizeof(REAL) * (order + 1));
        REAL pr = 1, pi = 0;
        for (int j = 1; j <= order / i; ++j) {
            const REAL tr = (pr * facr[i] - pi * faci[i]) / j, ti = (pr * faci[i] + pi * facr[i]) / j;
            pr = tr; pi = ti;
            for (int k = i * j; k <= order; ++k) {
                nr[k] += cr[k - i * j] * pr - ci[k - i * j] * pi;
                ni[k] += cr[k - i * j] * pi + ci[k - i * j] * pr;
            }
        }
        REAL* t = cr; cr = nr; nr = t; t = ci; ci = ni; ni = t;
    }
    if (cr != outr) { memcpy(outr, cr, sizeof(REAL) * (order + 1)); memcpy(outi, ci, sizeof(REAL) * (order + 1)); }
}

/* Kahan accumulator so that the oracle's own summation error does not pollute comparisons */
typedef struct { REAL s, c; } kahan_t;
static inline void kadd(kahan_t* k, REAL x) { REAL y = x - k->c; REAL t = k->s + y; k->c = (t - k->s) - y; k->s = t; }

/* A: n x n complex128 interleaved, matched order (vertex i paired with i + n/2); D: n complex or NULL.
 * Sum over Glynn subset indices [j0, j1) without the final 0.5^(m-1) scale.  out = {re, im}. */
static void hafnian_range_impl(const double* A, const double* D, int n, uint64_t j0, uint64_t j1, int nthreads, double* out) {
    const int m = n / 2, ld = NPAD(n);
    kahan_t tot_r = {0, 0}, tot_i = {0, 0};
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        const size_t msz = (size_t)n * ld;
        REAL* buf = (REAL*)calloc(6 * msz + 8 * (size_t)ld + 8 * (size_t)(m + 2), sizeof(REAL));
        REAL *Mr = buf, *Mi = Mr + msz, *Pr = Mi + msz, *Pi = Pr + msz, *Qr = Pi + msz, *Qi = Qr + msz;
        REAL *xdr = Qi + msz, *xdi = xdr + ld, *dr = xdi + ld, *di = dr + ld, *tr_ = di + ld, *ti_ = tr_ + ld;
        REAL *facr = ti_ + ld + 2 * ld, *faci = facr + (m + 2), *cr = faci + (m + 2), *ci = cr + (m + 2), *t1 = ci + (m + 2), *t2 = t1 + (m + 2);
        REAL *ptr_r = t2 + (m + 2), *ptr_i = ptr_r + (m + 2);
        kahan_t acc_r = {0, 0}, acc_i = {0, 0};
#pragma omp for schedule(dynamic, 4)
        for (uint64_t j = j0; j < j1; ++j) {
            /* delta_i = 2 kept_i - 1, kept = bits of j MSB first (find_kept_edges); AX = A X diag(delta,delta) */
            int nk = 0;
            for (int c = 0; c < n; ++c) {
                const int e = c % m;
                const int kept = (int)((j >> (m - 1 - e)) & 1ull);
                if (c < m) nk += kept;
                const REAL dlt = kept ? 1 : -1;
                const int sc = c < m ? c + m : c - m;
                for (int r = 0; r < n; ++r) {
                    Mr[(size_t)r * ld + c] = dlt * A[2 * ((size_t)r * n + sc)];
                    Mi[(size_t)r * ld + c] = dlt * A[2 * ((size_t)r * n + sc) + 1];
                }
                if (D) {
                    xdr[c] = dlt * D[2 * sc]; xdi[c] = dlt * D[2 * sc + 1];
                    dr[c] = D[2 * c]; di[c] = D[2 * c + 1];
                }
            }
            /* power traces tr(AX^k), k = 1..m, by the full product chain (charpoly.powertrace) */
            memcpy(Pr, Mr, sizeof(REAL) * msz); memcpy(Pi, Mi, sizeof(REAL) * msz);
            REAL *cPr = Pr, *cPi = Pi, *nPr = Qr, *nPi = Qi;
            for (int k = 1; k <= m; ++k) {
                if (k > 1) {
                    cmatmul(n, ld, cPr, cPi, Mr, Mi, nPr, nPi);
                    REAL* t = cPr; cPr = nPr; nPr = t; t = cPi; cPi = nPi; nPi = t;
                }
                REAL sr = 0, si = 0;
                for (int r = 0; r < n; ++r) { sr += cPr[(size_t)r * ld + r]; si += cPi[(size_t)r * ld + r]; }
                ptr_r[k] = sr; ptr_i[k] = si;
            }
            for (int i = 1; i <= m; ++i) {
                facr[i] = ptr_r[i] / (2 * i); faci[i] = ptr_i[i] / (2 * i);
                if (D) {
                    /* + (XD . D)/2 then XD <- XD AX  (f_loop) */
                    REAL sr = 0, si = 0;
                    for (int c = 0; c < n; ++c) { sr += xdr[c] * dr[c] - xdi[c] * di[c]; si += xdr[c] * di[c] + xdi[c] * dr[c]; }
                    facr[i] += sr / 2; faci[i] += si / 2;
                    for (int c = 0; c < n; ++c) { tr_[c] = 0; ti_[c] = 0; }
                    for (int r = 0; r < n; ++r)
                        for (int c = 0; c < n; ++c) {
                            tr_[c] += xdr[r] * Mr[(size_t)r * ld + c] - xdi[r] * Mi[(size_t)r * ld + c];
                            ti_[c] += xdr[r] * Mi[(size_t)r * ld + c] + xdi[r] * Mr[(size_t)r * ld + c];
                        }
                    memcpy(xdr, tr_, sizeof(REAL) * n); memcpy(xdi, ti_, sizeof(REAL) * n);
                }
            }
            exp_series(m, facr, faci, cr, ci, t1, t2);
            const REAL sg = ((m - nk) & 1) ? -1 : 1;
            kadd(&acc_r, sg * cr[m]); kadd(&acc_i, sg * ci[m]);
        }
#pragma omp critical
        { kadd(&tot_r, acc_r.s); kadd(&tot_r, -acc_r.c); kadd(&tot_i, acc_i.s); kadd(&tot_i, -acc_i.c); }
        free(buf);
    }
    out[0] = (double)tot_r.s; out[1] = (double)tot_i.s;
#ifdef ORACLE_LONG_DOUBLE
    out[2] = (double)(tot_r.s - (REAL)out[0]); out[3] = (double)(tot_i.s - (REAL)out[1]);
#else
    out[2] = 0; out[3] = 0;
#endif
}

void oracle_hafnian_range(const double* A, int n, uint64_t j0, uint64_t j1, int nthreads, double* out4) {
    hafnian_range_impl(A, NULL, n, j0, j1, nthreads, out4);
}
void oracle_loop_hafnian_range(const double* A, const double* D, int n, uint64_t j0, uint64_t j1, int nthreads, double* out4) {
    hafnian_range_impl(A, D, n, j0, j1, nthreads, out4);
}

/* Gray-code permanent over steps [k0, k1): method 0 = bbfg (no 2^(1-n) scale), 1 = ryser.  Serial per
 * thread like the reference; threads split the range (the reference itself is single-threaded). */
void oracle_perm_range(const double* M, int n, int method, uint64_t k0, uint64_t k1, int nthreads, double* out4) {
    kahan_t tot_r = {0, 0}, tot_i = {0, 0};
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        int tid = 0, nt = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
        const uint64_t len = k1 - k0;
        const uint64_t a = k0 + (len * tid) / nt, b = k0 + (len * (tid + 1)) / nt;
        REAL rr[64], ri[64];
        kahan_t acc_r = {0, 0}, acc_i = {0, 0};
        if (a < b) {
            uint64_t gray = a ^ (a >> 1);
            for (int c = 0; c < n; ++c) { rr[c] = 0; ri[c] = 0; }
            for (int r = 0; r < n; ++r) {
                const int set = (int)((gray >> r) & 1ull);
                const REAL d = method ? (set ? -1 : 0) : (set ? -1 : 1);
                for (int c = 0; c < n; ++c) { rr[c] += d * M[2 * ((size_t)r * n + c)]; ri[c] += d * M[2 * ((size_t)r * n + c) + 1]; }
            }
            const REAL step = method ? 1 : 2;
            for (uint64_t k = a; k < b; ++k) {
                REAL pr = rr[0], pi = ri[0];
                for (int c = 1; c < n; ++c) { const REAL t = pr * rr[c] - pi * ri[c]; pi = pr * ri[c] + pi * rr[c]; pr = t; }
                if (k & 1ull) { kadd(&acc_r, -pr); kadd(&acc_i, -pi); } else { kadd(&acc_r, pr); kadd(&acc_i, pi); }
                const uint64_t k1n = k + 1;
                const int row = __builtin_ctzll(k1n);
                if (row < n) {
                    const int set = (int)(((k1n ^ (k1n >> 1)) >> row) & 1ull);
                    const REAL d = set ? -step : step;
                    for (int c = 0; c < n; ++c) { rr[c] += d * M[2 * ((size_t)row * n + c)]; ri[c] += d * M[2 * ((size_t)row * n + c) + 1]; }
                }
            }
        }
#pragma omp critical
        { kadd(&tot_r, acc_r.s); kadd(&tot_r, -acc_r.c); kadd(&tot_i, acc_i.s); kadd(&tot_i, -acc_i.c); }
    }
    out4[0] = (double)tot_r.s; out4[1] = (double)tot_i.s;
#ifdef ORACLE_LONG_DOUBLE
    out4[2] = (double)(tot_r.s - (REAL)out4[0]); out4[3] = (double)(tot_i.s - (REAL)out4[1]);
#else
    out4[2] = 0; out4[3] = 0;
#endif
}

/* ---- torontonian ------------------------------------------------------------------------------ */
typedef struct { REAL re, im; } cplx;
static inline cplx cmulconj(cplx a, cplx b) { cplx r = {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im}; return r; }

/* Cholesky (lower) of the Hermitian matrix B (dim x dim, stride ld) starting from row `from`, rows < from
 * of L already valid.  Returns product of the diagonal from row `from` on. */
static REAL chol_from(const cplx* B, cplx* L, int dim, int ld, int from) {
    REAL prod = 1;
    for (int i = from; i < dim; ++i) {
        for (int j = from; j < i; ++j) {   /* columns < from are inherited (quad_cholesky :175-181) */
            cplx z = {0, 0};
            for (int k = 0; k < j; ++k) { cplx q = cmulconj(L[i * ld + k], L[j * ld + k]); z.re += q.re; z.im += q.im; }
            const REAL dj = L[j * ld + j].re;
            L[i * ld + j].re = (B[i * ld + j].re - z.re) / dj;
            L[i * ld + j].im = (B[i * ld + j].im - z.im) / dj;
        }
        REAL z = 0;
        for (int k = 0; k < i; ++k) z += L[i * ld + k].re * L[i * ld + k].re + L[i * ld + k].im * L[i * ld + k].im;
        L[i * ld + i].re = SQRT(B[i * ld + i].re - z);
        L[i * ld + i].im = 0;
        prod *= L[i * ld + i].re;
    }
    return prod;
}

/* recursiveTor (_torontonian.py:189-224): B = I - A interleaved, L its Cholesky factor; delete mode i >= start */
static REAL rec_tor(const cplx* B, const cplx* L, int nm, int ndel, int start, int n) {
    REAL tot = 0;
    const int dim = 2 * nm;
    if (nm == 0) return 0;
    const int cd = dim - 2;
    cplx* Bz = (cplx*)malloc(sizeof(cplx) * (size_t)(cd > 0 ? cd * cd : 1) * 2);
    cplx* Lz = Bz + (size_t)(cd > 0 ? cd * cd : 1);
    for (int i = start; i < n; ++i) {
        const int idx = (i - ndel) * 2;
        /* Z = all rows but idx, idx+1 */
        for (int r = 0, rr = 0; r < dim; ++r) {
            if (r == idx || r == idx + 1) continue;
            for (int c = 0, cc = 0; c < dim; ++c) {
                if (c == idx || c == idx + 1) continue;
                Bz[rr * cd + cc] = B[r * dim + c];
                Lz[rr * cd + cc] = L[r * dim + c];
                ++cc;
            }
            ++rr;
        }
        /* quad_cholesky: rows >= idx recomputed, earlier rows reused; det = (prod diag)^2 over ALL rows */
        chol_from(Bz, Lz, cd, cd, idx);
        REAL pd = 1;
        for (int r = 0; r < cd; ++r) pd *= Lz[r * cd + r].re;
        const REAL sign = ((ndel + 1) & 1) ? -1 : 1;
        tot += sign / pd + rec_tor(Bz, Lz, nm - 1, ndel + 1, i + 1, n);
    }
    free(Bz);
    return tot;
}

/* O: 2N x 2N complex interleaved doubles, block order.  Returns the torontonian (real). */
double oracle_tor_recursive(const double* O, int N) {
    const int dim = 2 * N;
    cplx* B = (cplx*)malloc(sizeof(cplx) * (size_t)dim * dim * 2);
    cplx* L = B + (size_t)dim * dim;
    for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
            const int sr = (r >> 1) + (r & 1) * N, sc = (c >> 1) + (c & 1) * N;
            B[r * dim + c].re = (r == c ? 1 : 0) - O[2 * ((size_t)sr * dim + sc)];
            B[r * dim + c].im = -O[2 * ((size_t)sr * dim + sc) + 1];
            L[r * dim + c].re = 0; L[r * dim + c].im = 0;
        }
    const REAL pd = chol_from(B, L, dim, dim, 0);
    const REAL res = 1 / pd + rec_tor(B, L, N, 0, 0, N);
    free(B);
    return (double)res;
}

/* numba_tor: flat loop over subsets [j0, j1), bit (N-1-i) of j <-> mode i kept */
double oracle_tor_direct_range(const double* O, int N, uint64_t j0, uint64_t j1, int nthreads) {
    kahan_t tot = {0, 0};
    const int dimO = 2 * N;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cplx* B = (cplx*)malloc(sizeof(cplx) * (size_t)dimO * dimO * 2);
        cplx* L = B + (size_t)dimO * dimO;
        int rows[128];
        kahan_t acc = {0, 0};
#pragma omp for schedule(dynamic, 64)
        for (uint64_t j = j0; j < j1; ++j) {
            int k = 0;
            for (int i = 0; i < N; ++i)
                if ((j >> (N - 1 - i)) & 1ull) { rows[2 * k] = i; rows[2 * k + 1] = i + N; ++k; }
            const int dim = 2 * k;
            for (int r = 0; r < dim; ++r)
                for (int c = 0; c < dim; ++c) {
                    B[r * dim + c].re = (r == c ? 1 : 0) - O[2 * ((size_t)rows[r] * dimO + rows[c])];
                    B[r * dim + c].im = -O[2 * ((size_t)rows[r] * dimO + rows[c]) + 1];
                }
            const REAL pd = dim ? chol_from(B, L, dim, dim, 0) : 1;
            kadd(&acc, (((N - k) & 1) ? -1 : 1) / pd);
        }
#pragma omp critical
        { kadd(&tot, acc.s); kadd(&tot, -acc.c); }
        free(B);
    }
    return (double)tot.s;
}

/* numba_ltor (thewalrus/_torontonian.py:369-412): flat loop over subsets [j0, j1) of
 * (-1)^(N-|S|) exp(gamma_S (I - O_S)^-1 gamma_S^* / 2) / sqrt(det(I - O_S)).  The quadratic form is evaluated
 * through the Cholesky factor, x^H B^-1 x = |L^-1 x|^2 with x = conj(gamma_S), as rec_ltorontonian does
 * (:338-343).  out2 = {re, 0}: real for Hermitian O. */
void oracle_ltor_direct_range(const double* O, const double* gamma, int N, uint64_t j0, uint64_t j1, int nthreads,
                              double* out2) {
    kahan_t tot = {0, 0};
    const int dimO = 2 * N;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cplx* B = (cplx*)malloc(sizeof(cplx) * ((size_t)dimO * dimO * 2 + dimO));
        cplx* L = B + (size_t)dimO * dimO;
        cplx* z = L + (size_t)dimO * dimO;
        int rows[128];
        kahan_t acc = {0, 0};
#pragma omp for schedule(dynamic, 64)
        for (uint64_t j = j0; j < j1; ++j) {
            int k = 0;
            for (int i = 0; i < N; ++i)
                if ((j >> (N - 1 - i)) & 1ull) { rows[k] = i; ++k; }
            for (int i = 0; i < k; ++i) rows[k + i] = rows[i] + N;
            const int dim = 2 * k;
            for (int r = 0; r < dim; ++r) {
                for (int c = 0; c < dim; ++c) {
                    B[r * dim + c].re = (r == c ? 1 : 0) - O[2 * ((size_t)rows[r] * dimO + rows[c])];
                    B[r * dim + c].im = -O[2 * ((size_t)rows[r] * dimO + rows[c]) + 1];
                }
                z[r].re = gamma[2 * rows[r]];
                z[r].im = -gamma[2 * rows[r] + 1];          /* x = conj(gamma_S) */
            }
            const REAL pd = dim ? chol_from(B, L, dim, dim, 0) : 1;
            REAL q = 0;
            for (int r = 0; r < dim; ++r) {                 /* forward substitution L z = x */
                cplx s = z[r];
                for (int c = 0; c < r; ++c) {
                    const cplx l = L[r * dim + c];
                    s.re -= l.re * z[c].re - l.im * z[c].im;
                    s.im -= l.re * z[c].im + l.im * z[c].re;
                }
                const REAL d = L[r * dim + r].re;
                z[r].re = s.re / d; z[r].im = s.im / d;
                q += z[r].re * z[r].re + z[r].im * z[r].im;
            }
#ifdef ORACLE_LONG_DOUBLE
            const REAL ex = expl(q / 2);
#else
            const REAL ex = exp(q / 2);
#endif
            kadd(&acc, (((N - k) & 1) ? -1 : 1) * ex / pd);
        }
#pragma omp critical
        { kadd(&tot, acc.s); kadd(&tot, -acc.c); }
        free(B);
    }
    out2[0] = (double)tot.s; out2[1] = 0;
}

/* brs / ubrs (thewalrus/_permanent.py:198-249): sum over row-subset labels [j0, j1) of A (m x n; bit (m-1-i) of the
 * label keeps row i) of (-1)^(m-|Y|) perm_bbfg(A_Y^H A_Y + E); E may be NULL (ubrs).  perm_bbfg includes its
 * 2^(1-n) factor (:167).  out2 = {re, im}. */
void oracle_brs_range(const double* A, const double* E, int m, int n, uint64_t j0, uint64_t j1, int nthreads,
                      double* out2) {
    kahan_t tot_r = {0, 0}, tot_i = {0, 0};
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        REAL* G = (REAL*)malloc(sizeof(REAL) * 2 * (size_t)n * n);
        REAL rr[64], ri[64];
        kahan_t acc_r = {0, 0}, acc_i = {0, 0};
#pragma omp for schedule(dynamic, 1)
        for (uint64_t j = j0; j < j1; ++j) {
            int cnt = 0;
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) {
                    G[2 * (a * n + b)] = E ? E[2 * ((size_t)a * n + b)] : 0;
                    G[2 * (a * n + b) + 1] = E ? E[2 * ((size_t)a * n + b) + 1] : 0;
                }
            for (int i = 0; i < m; ++i) {
                if (!((j >> (m - 1 - i)) & 1ull)) continue;
                ++cnt;
                for (int a = 0; a < n; ++a) {               /* G += conj(A[i, a]) A[i, b] */
                    const REAL ar = A[2 * ((size_t)i * n + a)], ai = -A[2 * ((size_t)i * n + a) + 1];
                    for (int b = 0; b < n; ++b) {
                        const REAL br = A[2 * ((size_t)i * n + b)], bi = A[2 * ((size_t)i * n + b) + 1];
                        G[2 * (a * n + b)] += ar * br - ai * bi;
                        G[2 * (a * n + b) + 1] += ar * bi + ai * br;
                    }
                }
            }
            /* perm_bbfg(G): Gray-code Glynn over 2^(n-1) steps */
            REAL pr_tot = 0, pi_tot = 0;
            for (int c = 0; c < n; ++c) { rr[c] = 0; ri[c] = 0; }
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) { rr[c] += G[2 * (r * n + c)]; ri[c] += G[2 * (r * n + c) + 1]; }
            const uint64_t steps = n > 0 ? (1ull << (n - 1)) : 1;
            for (uint64_t k = 0; k < steps; ++k) {
                REAL pr = 1, pi = 0;
                for (int c = 0; c < n; ++c) { const REAL t = pr * rr[c] - pi * ri[c]; pi = pr * ri[c] + pi * rr[c]; pr = t; }
                if (k & 1ull) { pr_tot -= pr; pi_tot -= pi; } else { pr_tot += pr; pi_tot += pi; }
                const uint64_t k1n = k + 1;
                const int row = __builtin_ctzll(k1n);
                if (row < n - 1 || (row < n && k1n < steps)) {
                    const int set = (int)(((k1n ^ (k1n >> 1)) >> row) & 1ull);
                    const REAL d = set ? -2 : 2;
                    for (int c = 0; c < n; ++c) { rr[c] += d * G[2 * (row * n + c)]; ri[c] += d * G[2 * (row * n + c) + 1]; }
                }
            }
            const REAL scale = (((m - cnt) & 1) ? -1 : 1) / (REAL)steps;
            kadd(&acc_r, scale * pr_tot); kadd(&acc_i, scale * pi_tot);
        }
#pragma omp critical
        { kadd(&tot_r, acc_r.s); kadd(&tot_r, -acc_r.c); kadd(&tot_i, acc_i.s); kadd(&tot_i, -acc_i.c); }
        free(G);
    }
    out2[0] = (double)tot_r.s; out2[1] = (double)tot_i.s;
}

/* ---- loop hafnian with repeated vertices (general path) ------------------------------------------- */
#ifdef ORACLE_LONG_DOUBLE
typedef long double _Complex zc;
#else
typedef double _Complex zc;
#endif
#define ZC(re, im) ((REAL)(re) + (REAL)(im) * (zc)_Complex_I)

/* matched_reps (_hafnian.py:80-159): greedy pairing by (count, index) descending.  Returns the number of edges. */
static int lh_matched_reps(const int32_t* reps, int nv, int* ea, int* eb, int* er, int* odd) {
    int cnt[128], live = 0, E = 0;
    for (int i = 0; i < nv; ++i) { cnt[i] = reps[i]; live += cnt[i] > 0; }
    for (;;) {
        int i0 = -1, i1 = -1;
        for (int i = nv - 1; i >= 0; --i) {       /* descending index: ties pick the larger index first */
            if (cnt[i] <= 0) continue;
            if (i0 < 0 || cnt[i] > cnt[i0]) { i1 = i0; i0 = i; }
            else if (i1 < 0 || cnt[i] > cnt[i1]) i1 = i;
        }
        if (live == 0 || (live == 1 && cnt[i0] <= 1)) break;
        if (live == 1 || cnt[i0] > 2 * cnt[i1]) {
            ea[E] = i0; eb[E] = i0; er[E] = cnt[i0] / 2;
            if (cnt[i0] & 1) cnt[i0] = 1; else { cnt[i0] = 0; --live; }
        } else {
            ea[E] = i0; eb[E] = i1; er[E] = cnt[i1];
            cnt[i0] -= cnt[i1]; cnt[i1] = 0; --live;
            if (cnt[i0] == 0) --live;
        }
        ++E;
    }
    *odd = -1;
    for (int i = 0; i < nv; ++i) if (cnt[i] > 0) *odd = i;
    return E;
}

static REAL lh_binom(int n, int k) {
    REAL b = 1;
    if (k > n - k) k = n - k;
    for (int q = 0; q < k; ++q) b = b * (REAL)(n - q) / (REAL)(q + 1);
    return b;
}

/* exp series of f / f_loop / f_loop_odd (`comb` loops, _hafnian.py:196-209, 228-242, 264-285) */
static zc lh_exp_series(const zc* fac, int order, zc* comb, zc* tmp) {
    for (int i = 0; i <= order; ++i) comb[i] = 0;
    comb[0] = 1;
    for (int i = 1; i <= order; ++i) {
        for (int q = 0; q <= order; ++q) tmp[q] = comb[q];
        zc pw = 1;
        for (int j = 1; j <= order / i; ++j) {
            pw = pw * fac[i] / (REAL)j;
            for (int q = i * j; q <= order; ++q) tmp[q] += comb[q - i * j] * pw;
        }
        for (int q = 0; q <= order; ++q) comb[q] = tmp[q];
    }
    return comb[order];
}

/* one loop hafnian: A nv x nv, D nv (or NULL: hafnian_repeated), reps nv */
static zc lh_one(const double* A, const double* D, int nv, const int32_t* reps, int glynn) {
    int N = 0;
    for (int i = 0; i < nv; ++i) N += reps[i];
    if (N == 0) return 1;
    if (!D && (N & 1)) return 0;
    if (D && N == 1) { for (int i = 0; i < nv; ++i) if (reps[i] == 1) return ZC(D[2 * i], D[2 * i + 1]); }
    int ea[64], eb[64], er[64], odd;
    const int E = lh_matched_reps(reps, nv, ea, eb, er, &odd);
    if (odd >= 0 && !D) return 0;
    const int has_odd = odd >= 0;
    const int order = has_odd ? N : N / 2, T = N / 2, smax = 2 * E;
    zc* buf = (zc*)malloc(sizeof(zc) * ((size_t)3 * smax * smax + 6 * (size_t)smax + 4 * (size_t)(order + 2) + (size_t)(T + 2)));
    zc *M = buf, *P = M + (size_t)smax * smax, *Q = P + (size_t)smax * smax;
    zc *xd = Q + (size_t)smax * smax, *dd = xd + smax, *d2 = dd + smax, *ov = d2 + smax, *xd2 = ov + smax, *spare = xd2 + smax;
    zc *fac = spare + smax, *comb = fac + order + 2, *tmp = comb + order + 2, *ptr = tmp + order + 2;
    (void)spare;
    uint64_t steps = 1;
    for (int e = 0; e < E; ++e) steps *= (uint64_t)((e == 0 && glynn && !has_odd) ? (er[0] + 2) / 2 : er[e] + 1);
    int kept[64], rows[128];
    REAL delta[128];
    zc H = 0;
    for (uint64_t j = 0; j < steps; ++j) {
        uint64_t num = j;
        for (int e = E - 1; e >= 0; --e) { kept[e] = (int)(num % (uint64_t)(er[e] + 1)); num /= (uint64_t)(er[e] + 1); }
        int k = 0, esum = 0, d0zero = 0;
        REAL w = 1;
        for (int e = 0; e < E; ++e) {
            esum += kept[e];
            w *= lh_binom(er[e], kept[e]);
            const int d = glynn ? 2 * kept[e] - er[e] : kept[e];
            if (e == 0) d0zero = (d == 0);
            if (d != 0) { rows[k] = e; delta[k] = (REAL)d; ++k; }
        }
        for (int a = 0; a < k; ++a) { const int e = rows[a]; rows[a] = ea[e]; rows[k + a] = eb[e]; delta[k + a] = delta[a]; }
        const int s = 2 * k;
        /* get_submatrices (:315-356): M[:, c] = delta_c * A[rows, rows][:, swap(c)] */
        for (int r = 0; r < s; ++r)
            for (int c = 0; c < s; ++c) {
                const int sc = c < k ? c + k : c - k;
                const size_t at = 2 * ((size_t)rows[r] * nv + rows[sc]);
                M[r * s + c] = delta[c] * ZC(A[at], A[at + 1]);
            }
        for (int c = 0; c < s; ++c) {
            const int sc = c < k ? c + k : c - k;
            if (D) { xd[c] = delta[c] * ZC(D[2 * rows[sc]], D[2 * rows[sc] + 1]); dd[c] = ZC(D[2 * rows[c]], D[2 * rows[c] + 1]); }
            if (has_odd) { const size_t at = 2 * ((size_t)odd * nv + rows[sc]); ov[c] = delta[c] * ZC(A[at], A[at + 1]); }
        }
        /* power traces tr(M^i), i = 1..T, by the product chain (charpoly.py:316-318) */
        for (int i = 0; i < s * s; ++i) P[i] = M[i];
        for (int i = 1; i <= T; ++i) {
            zc tr = 0;
            for (int r = 0; r < s; ++r) tr += P[r * s + r];
            ptr[i] = tr;
            if (i < T) {
                for (int r = 0; r < s; ++r)
                    for (int c = 0; c < s; ++c) {
                        zc a = 0;
                        for (int q = 0; q < s; ++q) a += P[r * s + q] * M[q * s + c];
                        Q[r * s + c] = a;
                    }
                zc* t2 = P; P = Q; Q = t2;
            }
        }
        REAL pre = (((N / 2 - esum) & 1) ? -1 : 1) * w;
        if (!has_odd) {                                   /* f / f_loop (:183-242) */
            for (int c = 0; c < s; ++c) xd2[c] = D ? xd[c] : 0;
            for (int i = 1; i <= order; ++i) {
                zc l = 0;
                if (D) {
                    for (int c = 0; c < s; ++c) l += xd2[c] * dd[c];
                    for (int c = 0; c < s; ++c) { zc a = 0; for (int q = 0; q < s; ++q) a += xd2[q] * M[q * s + c]; d2[c] = a; }
                    for (int c = 0; c < s; ++c) xd2[c] = d2[c];
                }
                fac[i] = ptr[i] / (REAL)(2 * i) + l / 2;
            }
            if (glynn && d0zero) pre *= (REAL)0.5;
        } else {                                          /* f_loop_odd (:246-285) */
            for (int i = 1; i <= order; ++i) {
                if (i == 1) fac[i] = ZC(D[2 * odd], D[2 * odd + 1]);
                else if ((i & 1) == 0) {
                    zc l = 0;
                    for (int c = 0; c < s; ++c) l += xd[c] * dd[c];
                    fac[i] = ptr[i / 2] / (REAL)i + l / 2;
                } else {
                    zc o = 0;
                    for (int c = 0; c < s; ++c) o += ov[c] * dd[c];
                    fac[i] = o;
                    for (int r = 0; r < s; ++r) { zc a = 0; for (int q = 0; q < s; ++q) a += M[r * s + q] * dd[q]; d2[r] = a; }
                    for (int r = 0; r < s; ++r) dd[r] = d2[r];
                }
            }
        }
        H += pre * lh_exp_series(fac, order, comb, tmp);
    }
    if (glynn) H *= (REAL)ldexp(1.0, -(has_odd ? N / 2 : N / 2 - 1));
    free(buf);
    return H;
}

/* out[b] = loop_hafnian(A, D, reps = rpt[b]) (hafnian_repeated when D is NULL) for b < B; patterns in parallel */
void oracle_lhaf_patterns(const double* A, const double* D, int nv, const int32_t* rpt, int64_t B, int glynn,
                          int nthreads, double* out) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        const zc v = lh_one(A, D, nv, rpt + (size_t)b * nv, glynn);
        out[2 * b] = (double)creall(v);
        out[2 * b + 1] = (double)cimagl(v);
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
