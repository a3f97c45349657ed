/* walrus_b200 — C ABI of the B200-native exponential-sum hot path of The Walrus.
 *
 * The reference (XanaduAI/thewalrus v0.23.0-dev) is Python + Numba and has no FFI layer; these entry
 * points are what a ctypes binding placed at the reference's numba call sites would bind.  Each entry
 * cites the reference driver it replaces.  All matrices are row-major complex128 stored as interleaved
 * (re, im) doubles and are ALREADY in "matched" order (vertex i paired with i + n/2), exactly what the
 * reference passes to its numba drivers after matched_reps (thewalrus/_hafnian.py:501-505, 617-628).
 *
 * Every sum entry point evaluates a half-open range [j0, j1) of the reference's subset index, so the same
 * symbol serves single-GPU calls and contiguous multi-GPU shards; partial results come back as
 * compensated (hi, lo) pairs: out[0..3] = {re_hi, re_lo, im_hi, im_lo}, value = hi + lo, WITHOUT the final
 * power-of-two scaling (0.5^(N/2-1), 2^(1-n), ...), which the caller applies after combining shards.
 *
 * Return value: 0 on success, negative WB200_E* code otherwise (wb200_last_error() gives the text).
 * No C++ exceptions cross the boundary. `*_dev` variants take device pointers and a cudaStream_t
 * (as void*) and do not synchronise; `*_host` variants take host pointers, copy in/out and synchronise.
 */
#ifndef WALRUS_B200_H
#define WALRUS_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define WB200_OK 0
#define WB200_EINVAL -1   /* bad argument (size, parity, null pointer) */
#define WB200_ECUDA -2    /* CUDA runtime error, see wb200_last_error() */
#define WB200_ENOSUP -3   /* size outside what the kernels are built for */

const char* wb200_last_error(void);
int wb200_version(void);
int wb200_device_count(int* count);
/* The *_host wrappers keep their device scratch in a thread-local cache between calls (a wrapper called once per
 * sampler mode or per pattern batch would otherwise spend more time in cudaMalloc/cudaFree than in its kernel).
 * This returns the calling thread's idle blocks to the driver. */
int wb200_release_scratch(void);

/* FP64 pipe micro-benchmark used as the roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
 * kind 0 = DFMA chains, 1 = DMMA m8n8k4 chains.  Result in TFLOP/s. */
int wb200_fp64_peak(int device, int kind, double* tflops);

/* ---- hafnian / loop hafnian, all edge repetitions 1, Glynn sieve (fast DMMA path) -------------------
 * Replaces _calc_hafnian (thewalrus/_hafnian.py:416-467) and _calc_loop_hafnian (:512-577) for
 * edge_reps = [1]*m, glynn=True, no odd vertex.  n even, 2 <= n <= 64.  Subset index j in
 * [0, 2^(n/2-1)) as in :433/:536.  D = NULL: hafnian; D != NULL (n complex): loop hafnian.
 * Final scale (caller): 0.5^(n/2-1).
 * Two kernels serve this entry: the row-panel kernel (hafnian_dmma.cu, every size) and, for D = NULL, even n in
 * [36, 64] and ranges of at least 8 groups of four subsets per SM, the symmetric-half kernel
 * (hafnian_sym.cu: only the tiles on and above the diagonal of every product, 1.3 - 1.5 x faster).  Environment:
 * WB200_HAF_SYM=0 forces the row-panel kernel, =4 the one-team shape of the symmetric-half kernel (n = 48 / 50). */
size_t wb200_hafnian_workspace_bytes(int n);
int wb200_hafnian_dev(const double* dA, const double* dD, int n, uint64_t j0, uint64_t j1,
                      double* d_out4, void* d_workspace, size_t workspace_bytes, void* stream);
int wb200_hafnian_host(int device, const double* A, const double* D, int n, uint64_t j0, uint64_t j1,
                       double out4[4], double* kernel_ms);

/* ---- general (repeated edges, inclusion/exclusion, odd vertex) loop hafnian ------------------------
 * Replaces _calc_hafnian / _calc_loop_hafnian for arbitrary edge_reps (thewalrus/_hafnian.py:416-577):
 * mixed-radix subset index (find_kept_edges :162-180), binomial weights (:447-449), delta = 0 row/column
 * deletion (get_submatrices :315-356), f / f_loop / f_loop_odd (:183-285).
 * A: n x n (n = 2*n_edges) complex; D: n complex or NULL (no loops); oddV: n complex and oddloop: 2
 * doubles, or NULL when there is no unpaired vertex; edge_reps: n_edges int32.
 * out4 excludes the final 0.5^(N/2-1) (or 0.5^(N/2) with an odd vertex) Glynn scale. */
int wb200_lhaf_general_host(int device, const double* A, const double* D, const double* oddV,
                            const double* oddloop, int n, const int32_t* edge_reps, int glynn,
                            uint64_t j0, uint64_t j1, double out4[4], double* kernel_ms);
/* Device-pointer twin: dA, dD, doddV live on the current device; oddloop (2 doubles) and edge_reps are small HOST
 * arrays (launch metadata).  Asynchronous on `stream`; scratch is stream-ordered (cudaMallocAsync). */
int wb200_lhaf_general_dev(const double* dA, const double* dD, const double* doddV, const double* oddloop, int n,
                           const int32_t* edge_reps, int glynn, uint64_t j0, uint64_t j1, double* d_out4, void* stream);
/* total number of subset indices for these arguments (reference `steps`, _hafnian.py:432-435, 535-538) */
int wb200_lhaf_general_steps(const int32_t* edge_reps, int n_edges, int glynn, int has_odd, uint64_t* steps);

/* ---- batched front end: loop hafnians of many repetition patterns of ONE matrix ----------------------
 * Replaces the per-pattern Python loop of probabilities() / density_matrix_element
 * (thewalrus/quantum/fock_tensors.py:191-232, 412-421): for every row rpt[b] (nv repetition counts) this
 * evaluates loop_hafnian(A, D = gamma, reps = rpt[b]) (thewalrus/_hafnian.py:581-631), or
 * hafnian_repeated(A, rpt[b]) when gamma is NULL (:470-508), including the N = 0 / N = 1 / odd-N early exits
 * and the final Glynn scale.  The vertex pairing (matched_reps, :80-159) is done on the device.
 * A: nv x nv complex (NOT permuted), gamma: nv complex or NULL, rpt: B x nv int32, out: B complex (re, im).
 * nv <= 64; per pattern at most 32 edges and series order <= 200. */
int wb200_lhaf_patterns_host(int device, const double* A, const double* gamma, int nv, const int32_t* rpt,
                             int64_t B, int glynn, double* out, double* kernel_ms);
/* Same, with a table of loop vectors: gamma is n_gamma x nv complex and pattern b uses row gamma_index[b]
 * (gamma_index may be NULL when n_gamma == 1).  This is the call the batched chain-rule samplers make: one mode
 * step of generate_hafnian_sample / generate_torontonian_sample (thewalrus/samples.py:245-249, 459-470) for
 * ALL samples (and fan-out channels) at once, each with its own heterodyne-shifted gamma. */
int wb200_lhaf_patterns_multi_host(int device, const double* A, const double* gamma, int n_gamma,
                                   const int32_t* gamma_index, int nv, const int32_t* rpt, int64_t B, int glynn,
                                   double* out, double* kernel_ms);
/* The batched-MATRIX front end: additionally a table of matrices, A is n_A x nv x nv and problem b uses matrix
 * A_index[b] (NULL when n_A == 1).  One call evaluates B independent (loop) hafnians
 * loop_hafnian(A[A_index[b]], gamma[gamma_index[b]], reps = rpt[b]) — e.g. hafnian(A_b) for a stack of small
 * matrices with rpt = 1 (thewalrus/_hafnian.py:718-861 called in a Python loop by the reference's users). */
int wb200_lhaf_matrices_host(int device, const double* A, int n_A, const int32_t* A_index, const double* gamma,
                             int n_gamma, const int32_t* gamma_index, int nv, const int32_t* rpt, int64_t B, int glynn,
                             double* out, double* kernel_ms);

/* Device-pointer twin of the batched front end: every array (dA, dA_index, dgamma, dgamma_index, drpt, d_out) lives on
 * the current device.  The launches go to `stream`; the call synchronises that stream ONCE in the middle (the per-class
 * work totals computed by the prep kernel decide the launch shapes) and returns without waiting for the kernels.
 * Index tables are validated on the device (WB200_EINVAL if an entry is outside its table). */
int wb200_lhaf_matrices_dev(const double* dA, int n_A, const int32_t* dA_index, const double* dgamma, int n_gamma,
                            const int32_t* dgamma_index, int nv, const int32_t* drpt, int64_t B, int glynn,
                            double* d_out, void* stream);

/* ---- loop_hafnian_batch sweep -------------------------------------------------------------------------
 * Replaces _calc_loop_hafnian_batch_even / _odd (thewalrus/loop_hafnian_batch.py:51-208).  Ax (n x n, n = 2E),
 * Dx: already edge-ordered by add_batch_edges_even/odd (:211-257); edge_reps = [batch_max, fixed...] (even
 * variant) or [batch_max, 1, fixed...] (odd variant).  Subset index j in [0, prod(edge_reps + 1)) (:79, :155).
 * out: length x {re_hi, re_lo, im_hi, im_lo}, length = 2 batch_max + cutoff_extra + 1 (+1 for the odd
 * variant), WITHOUT the final 0.5^((N_fixed + j) / 2) scaling (:118-121), so shards can be added first. */
int wb200_lhaf_batch_steps(const int32_t* edge_reps, int n_edges, uint64_t* steps);
int wb200_lhaf_batch_host(int device, const double* Ax, const double* Dx, int n, const int32_t* edge_reps,
                          int odd_variant, int cutoff_extra, int glynn, uint64_t j0, uint64_t j1, double* out,
                          int length, double* kernel_ms);

/* Same sweep for n_D loop vectors at once: replaces _calc_loop_hafnian_batch_gamma_even / _odd
 * (thewalrus/loop_hafnian_batch_gamma.py:52-220).  Dx: n_D x n (row k = the edge-ordered loop vector of
 * displacement k); out: n_D x length x {re_hi, re_lo, im_hi, im_lo}, unscaled as above.  The reduced matrix and
 * its power traces are computed once per subset and shared by all n_D vectors. */
int wb200_lhaf_batch_gamma_host(int device, const double* Ax, const double* Dx, int n, int n_D,
                                const int32_t* edge_reps, int odd_variant, int cutoff_extra, int glynn,
                                uint64_t j0, uint64_t j1, double* out, int length, double* kernel_ms);

/* Device-pointer twin (dAx, dDx, d_out on the current device; edge_reps on the host); asynchronous on `stream`. */
int wb200_lhaf_batch_gamma_dev(const double* dAx, const double* dDx, int n, int n_D, const int32_t* edge_reps,
                               int odd_variant, int cutoff_extra, int glynn, uint64_t j0, uint64_t j1, double* d_out,
                               int length, void* stream);

/* ---- chain-rule photon-number sampler: all mode steps on the device ----------------------------------------
 * Replaces the per-mode loop of generate_hafnian_sample (thewalrus/samples.py:204-261) for S chains at once.  Inputs (host,
 * prepared as the reference does, :228-247): B = Amat(T)[:M, :M] (M x M complex), gamma0 = conj(pure_alpha) +
 * (het_alpha - pure_alpha) B^T (S x M complex), het = het_alpha (S x M complex), uniforms (M x S doubles in [0, 1): row i
 * holds the draws of mode step i).  Mode step i: gamma <- gamma - het[:, i] B[:, i]; p(k) ~ |lhaf(B[:i+1, :i+1],
 * gamma[:i+1], reps = (n_0 .. n_(i-1), k))|^2 / k!, k = 0 .. cutoff; n_i = the outcome numpy.random.choice returns for
 * that uniform.  det_out: S x M int32 photon numbers (rejection of chains — last mode at the cutoff, too many photons —
 * is the caller's, :253-261).  M <= 64, cutoff <= 63.  WB200_EINVAL if a probability row is NaN / negative / all zero. */
int wb200_hafnian_chains_host(int device, const double* B, const double* gamma0, const double* het, const double* uniforms,
                              int M, int64_t S, int cutoff, int32_t* det_out, double* kernel_ms);

/* ---- montrealer ---------------------------------------------------------------------------------------
 * Replaces montrealer / lmontrealer (thewalrus/_montrealer.py:37-102) as called by mtl / lmtl (:105-135):
 * A: 2n x 2n complex (the matrix passed to mtl, NOT pre-multiplied by Xmat), zeta: 2n complex or NULL.
 * Subset labels p in [p0, p1) of [0, 2^n) (label 0, the empty set, contributes nothing).
 * out8 = V{re_hi, re_lo, im_hi, im_lo}, W{...} with V = sum_p (-1)^(|p|+1) tr(Sigma_p^n) and
 * W = sum_p (-1)^(|p|+1) conj(zeta_p) Sigma_p^(n-1) zeta_p; the caller forms (-1)^(n+1) (V / 2n + W / 2). */
int wb200_mtl_host(int device, const double* A, const double* zeta, int n_modes, uint64_t p0, uint64_t p1,
                   double out8[8], double* kernel_ms);
/* Device-pointer twin; asynchronous on `stream`. */
int wb200_mtl_dev(const double* dA, const double* dzeta, int n_modes, uint64_t p0, uint64_t p1, double* d_out8,
                  void* stream);

/* ---- permanent --------------------------------------------------------------------------------------
 * Replaces perm_bbfg (thewalrus/_permanent.py:130-168; method 0, steps k in [0, 2^(n-1)), final scale
 * 2^(1-n)) and perm_ryser (:86-127; method 1, steps k in [0, 2^n), no scale).  Step k evaluates the
 * Gray code g(k) = k ^ (k >> 1) with sign (-1)^k, as the reference's loop does.  M: n x n complex,
 * 1 <= n <= 64 (one thread holds all column sums up to n = 40; four lanes share them above). */
int wb200_perm_dev(const double* dM, int n, int method, uint64_t k0, uint64_t k1, double* d_out4,
                   void* d_workspace, size_t workspace_bytes, void* stream);
size_t wb200_perm_workspace_bytes(int n);
int wb200_perm_host(int device, const double* M, int n, int method, uint64_t k0, uint64_t k1,
                    double out4[4], double* kernel_ms);
/* real (float64) matrix: same sweep in real arithmetic (the reference keeps the input dtype).  out2 = {hi, lo}. */
int wb200_perm_f64_host(int device, const double* M, int n, int method, uint64_t k0, uint64_t k1,
                        double out2[2], double* kernel_ms);
/* exact int64 permanent (numba specialises perm_* on int64 input and sums in wrapping int64).
 * out = sum over [k0,k1) WITHOUT the bbfg division by 2^(n-1). */
int wb200_perm_int64_host(int device, const int64_t* M, int n, int method, uint64_t k0, uint64_t k1,
                          int64_t* out, double* kernel_ms);

/* ---- Bristolian ------------------------------------------------------------------------------------
 * Replaces brs / ubrs (thewalrus/_permanent.py:198-249): sum over row subsets Y of A (m x n complex; label bit i,
 * MSB first, keeps row i) of (-1)^(m-|Y|) perm_bbfg(A_Y^H A_Y + E), labels j in [j0, j1) of [0, 2^m).
 * E: n x n complex or NULL (ubrs: E = NULL and j0 = 1).  out4 is the (hi, lo) sum WITHOUT the 2^(1-n) factor of
 * perm_bbfg (:167), which the caller applies.  m <= 40, n <= 32. */
int wb200_brs_host(int device, const double* A, const double* E, int m, int n, uint64_t j0, uint64_t j1,
                   double out4[4], double* kernel_ms);
/* Device-pointer twin; asynchronous on `stream`. */
int wb200_brs_dev(const double* dA, const double* dE, int m, int n, uint64_t j0, uint64_t j1, double* d_out4,
                  void* stream);

/* ---- torontonian ------------------------------------------------------------------------------------
 * Replaces rec_torontonian / numba_tor (thewalrus/_torontonian.py:123-154, 189-247):
 * sum over S subset of [N] of (-1)^(N-|S|) / sqrt(det(I - O_S)).  O: 2N x 2N complex Hermitian in the
 * reference's (x..., p...) block order (mode i <-> rows i and i+N).  The subset space is cut in
 * 2^P "prefixes" (choices for the first P modes); this call evaluates prefixes [p0, p1).
 * wb200_tor_num_prefixes gives 2^P for N.  out2 = {hi, lo} (the sum is real). */
int wb200_tor_num_prefixes(int n_modes, uint64_t* count);
size_t wb200_tor_workspace_bytes(int n_modes);
/* device pointers, no synchronisation; d_out4 = {hi, lo, 0, 0} */
int wb200_tor_dev(const double* dO, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                  void* d_workspace, size_t workspace_bytes, void* stream);
int wb200_tor_host(int device, const double* O, int n_modes, uint64_t p0, uint64_t p1, double out2[2],
                   double* kernel_ms);
/* the same for a REAL symmetric O (2N x 2N doubles, not interleaved complex): the subset tree runs in real arithmetic
 * (numba specialises the reference's njit kernels on the dtype the same way, thewalrus/_torontonian.py:123, 157, 189).
 * Same prefixes, same workspace size, same outputs. */
int wb200_tor_f64_dev(const double* dO, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                      void* d_workspace, size_t workspace_bytes, void* stream);
int wb200_tor_f64_host(int device, const double* O, int n_modes, uint64_t p0, uint64_t p1, double out2[2],
                       double* kernel_ms);

/* ---- loop torontonian -------------------------------------------------------------------------------
 * Replaces rec_ltorontonian / recursiveLTor / numba_ltor (thewalrus/_torontonian.py:276-345, 369-412):
 * sum over S of (-1)^(N-|S|) exp(gamma_S (I - O_S)^-1 gamma_S^* / 2) / sqrt(det(I - O_S)).
 * O as for wb200_tor_*; gamma: 2N complex in the same (x..., p...) order.  Same prefix ranges, workspace and
 * output convention as the torontonian (the sum is real for Hermitian O, as in rec_ltorontonian). */
int wb200_ltor_dev(const double* dO, const double* dGamma, int n_modes, uint64_t p0, uint64_t p1, double* d_out4,
                   void* d_workspace, size_t workspace_bytes, void* stream);
int wb200_ltor_host(int device, const double* O, const double* gamma, int n_modes, uint64_t p0, uint64_t p1,
                    double out2[2], double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif
