// Batched loop hafnians of repetition patterns on the FP64 tensor cores (included by batch.cu).
//
// A pattern (thewalrus/quantum/fock_tensors.py:191-232 -> loop_hafnian(A, gamma, reps), _hafnian.py:512-577) with E
// matched edges is a mixed-radix subset sum over j; every subset uses the SAME gathered 2E x 2E matrix
// A'' = A[verts, verts] and differs only in the integer vector delta (delta_e = 2 kept_e - r_e).  With
// S_j = X diag(delta, delta) (symmetric, it commutes with the pair swap X) the reduced matrix of the reference is
// M_j = A'' S_j with the delta = 0 rows/columns DELETED (get_submatrices, _hafnian.py:315-356); zeroing them instead
// leaves every tr(M^k) and every loop term unchanged (M becomes block triangular).  So the row-panel machinery of
// hafnian_dmma.cu applies verbatim with the sign flip replaced by a multiplication by delta:
//     B_k := M^(k-1) A'' (symmetric),  row r of B_(k+1) = (row r of B_k) S A'',
//     tr(M^k)     = sum_r delta_r B_k[r, sigma(r)],
//     tr(M^(a+b)) = sum_r delta_r < (row r of B_a) S, row sigma(r) of B_b >,
//     l_(a+b+1)   = XD^T M^(a+b) D = < z_a S, z_b >,  z_k = D^T (S A'')^k.
// A warp owns a group of 4 consecutive subsets of one pattern (8 rows = 4 subsets x the vertex pair {e, e + E}) and
// walks the E row panels (+ the loop row); the B-operand fragment table of A'' is built per pattern in the warp's own
// shared-memory slice (a few thousand cycles against ~10^5 per chunk), so the warps of a CTA work on different
// patterns with no CTA-level synchronisation.  Work items are (pattern, chunk of BW_CHUNK subsets) pulled from an
// atomic counter; one compensated partial per chunk, summed per pattern in index order by pat_final_kernel.
//
// Patterns are bucketed by tile shape (TF full tiles of four vertex pairs + an optional packed tail tile): seven
// template instances cover E <= 16; odd totals, E > 16 and series orders > PD_TMAX stay on the warp-per-subset DFMA
// kernel (pat_main_kernel).
#pragma once
#include "haf_dmma.cuh"

namespace wb {

constexpr int PD_TMAX = 24;      // series order (N / 2) limit of the DMMA path
constexpr int PD_EMAX = 16;      // edges
constexpr int PD_WARPS = 12;     // warps per CTA, one CTA per SM (as haf_dmma_kernel); fewer where the per-warp tables are large
constexpr int PD_NCLS = 7;       // tile-shape classes; class PD_NCLS = fallback (pat_main_kernel)

// tile-shape class of a pattern with E edges: index into {(1,0),(1,1),(2,0),(2,1),(3,0),(3,1),(4,0)}
__host__ __device__ inline int pd_class(int E) {
    if (E < 1 || E > PD_EMAX) return PD_NCLS;
    int TF, tail;
    const int r = E & 3;
    if (E >= 5 && (r == 1 || r == 2)) { TF = E >> 2; tail = 1; }
    else { TF = (E + 3) >> 2; tail = 0; }
    return (TF - 1) * 2 + tail;
}

template <int TF, bool TAIL>
struct PatCfg {
    static constexpr int NK = 2 * TF + (TAIL ? 1 : 0);
    static constexpr int NT = TF + (TAIL ? 1 : 0);
    static constexpr int FRAG_D = NK * NT * 64;             // doubles in the fragment table
    static constexpr int PART_D = (PD_TMAX + 2) * 8 * 2;    // part[j][g] complex
    static constexpr int P_D = (PD_TMAX + 2) * 4 * 2;       // P[k][q] complex (power traces; loop terms; odd-row terms)
    static constexpr int C_D = P_D;                         // series factors / coefficients: order <= PD_TMAX (odd totals: N <= PD_TMAX)
    static constexpr int DELTA_D = 4 * PD_EMAX;             // delta[q][e]
    static constexpr int INFO_D = 4 + 2 * PD_EMAX + PD_EMAX + PD_EMAX / 2;   // pre[4], stride[16] (u64), verts[32] int, reps[16] int
    static constexpr int WARP_D = FRAG_D + PART_D + 3 * P_D + C_D + DELTA_D + INFO_D;   // the factors F reuse part[]
    // every warp carries its own pattern's fragment table: as many warps as fit in 227 KB, at most PD_WARPS
    static constexpr int FIT = (227 * 1024 - 1024) / (int)(sizeof(double) * WARP_D);
    static constexpr int WARPS = FIT < PD_WARPS ? FIT : PD_WARPS;
    static constexpr size_t BYTES = sizeof(double) * (size_t)WARPS * WARP_D;
};

// Y <- (W with partners swapped) * delta.  If IP: odd = <X, Y_old>, even = <X, Y_new>, X = W of the partner row
// (lane ^ 16) or W itself (SELF, the loop row).  Same contract as haf_advance with doubles instead of sign masks.
template <int TF, bool TAIL, bool IP, bool SELF>
__device__ __forceinline__ void pat_advance(const HafRow<TF, TAIL>& w, HafY<TF, TAIL>& y, const double* dl, double dlt,
                                            double& orr, double& oi, double& er, double& ei) {
    double o2r = 0.0, o2i = 0.0, e2r = 0.0, e2i = 0.0;
    orr = oi = er = ei = 0.0;
#pragma unroll
    for (int tau = 0; tau < TF; ++tau) {
        double x0r = 0, x0i = 0, x1r = 0, x1i = 0;
        if (IP) {
            x0r = SELF ? w.wr[tau][0] : shfl_xor_d(w.wr[tau][0], 16);
            x0i = SELF ? w.wi[tau][0] : shfl_xor_d(w.wi[tau][0], 16);
            x1r = SELF ? w.wr[tau][1] : shfl_xor_d(w.wr[tau][1], 16);
            x1i = SELF ? w.wi[tau][1] : shfl_xor_d(w.wi[tau][1], 16);
            WB_CFMA(orr, oi, x0r, x0i, y.yr[2 * tau], y.yi[2 * tau]);
            WB_CFMA(o2r, o2i, x1r, x1i, y.yr[2 * tau + 1], y.yi[2 * tau + 1]);
        }
        y.yr[2 * tau + 0] = dl[tau] * w.wr[tau][1];
        y.yr[2 * tau + 1] = dl[tau] * w.wr[tau][0];
        y.yi[2 * tau + 0] = dl[tau] * w.wi[tau][1];
        y.yi[2 * tau + 1] = dl[tau] * w.wi[tau][0];
        if (IP) {
            WB_CFMA(er, ei, x0r, x0i, y.yr[2 * tau], y.yi[2 * tau]);
            WB_CFMA(e2r, e2i, x1r, x1i, y.yr[2 * tau + 1], y.yi[2 * tau + 1]);
        }
    }
    if (TAIL) {
        double xr = 0, xi = 0;
        if (IP) {
            xr = SELF ? w.wtr : shfl_xor_d(w.wtr, 16);
            xi = SELF ? w.wti : shfl_xor_d(w.wti, 16);
            WB_CFMA(orr, oi, xr, xi, y.ytr, y.yti);
        }
        y.ytr = dlt * shfl_xor_d(w.wtr, 1);
        y.yti = dlt * shfl_xor_d(w.wti, 1);
        if (IP) WB_CFMA(er, ei, xr, xi, y.ytr, y.yti);
    }
    if (IP) {
        orr += o2r; oi += o2i; er += e2r; ei += e2i;
        orr += shfl_xor_d(orr, 1); oi += shfl_xor_d(oi, 1); er += shfl_xor_d(er, 1); ei += shfl_xor_d(ei, 1);
        orr += shfl_xor_d(orr, 2); oi += shfl_xor_d(oi, 2); er += shfl_xor_d(er, 2); ei += shfl_xor_d(ei, 2);
    }
}

// <X, Y> over this thread's slots, reduced over the four lanes of the row; X = W itself (SELF) or W of the partner
// row (lane ^ 16).  Used by the loop row of patterns with an unpaired vertex, where half 0 carries the chain of D and
// half 1 the chain of the odd vertex's row of A.
template <int TF, bool TAIL, bool SELF>
__device__ __forceinline__ void pat_inner(const HafRow<TF, TAIL>& w, const HafY<TF, TAIL>& y, double& sr, double& si) {
    double s2r = 0.0, s2i = 0.0;
    sr = si = 0.0;
#pragma unroll
    for (int tau = 0; tau < TF; ++tau) {
        const double x0r = SELF ? w.wr[tau][0] : shfl_xor_d(w.wr[tau][0], 16), x0i = SELF ? w.wi[tau][0] : shfl_xor_d(w.wi[tau][0], 16);
        const double x1r = SELF ? w.wr[tau][1] : shfl_xor_d(w.wr[tau][1], 16), x1i = SELF ? w.wi[tau][1] : shfl_xor_d(w.wi[tau][1], 16);
        WB_CFMA(sr, si, x0r, x0i, y.yr[2 * tau], y.yi[2 * tau]);
        WB_CFMA(s2r, s2i, x1r, x1i, y.yr[2 * tau + 1], y.yi[2 * tau + 1]);
    }
    if (TAIL) {
        const double xr = SELF ? w.wtr : shfl_xor_d(w.wtr, 16), xi = SELF ? w.wti : shfl_xor_d(w.wti, 16);
        WB_CFMA(sr, si, xr, xi, y.ytr, y.yti);
    }
    sr += s2r; si += s2i;
    sr += shfl_xor_d(sr, 1); si += shfl_xor_d(si, 1);
    sr += shfl_xor_d(sr, 2); si += shfl_xor_d(si, 2);
}

template <int TF, bool TAIL>
__global__ void __launch_bounds__(32 * PatCfg<TF, TAIL>::WARPS, 1) pat_dmma_kernel(PatParams p) {
    using C = PatCfg<TF, TAIL>;
    extern __shared__ __align__(16) double smem_pd[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3, q = g & 3, half = g >> 2;
    double* wsm = smem_pd + (size_t)warp * C::WARP_D;
    double2* sfrag = reinterpret_cast<double2*>(wsm);
    double2* part = reinterpret_cast<double2*>(wsm + C::FRAG_D);     // part[j * 8 + g]: row-g share of tr(M^j)
    double* Pk = wsm + C::FRAG_D + C::PART_D;                         // P[k][q] complex
    double* Lk = Pk + C::P_D;                                         // loop terms l_t
    double* Ok = Lk + C::P_D;                                         // odd-row terms o_t (patterns with an unpaired vertex)
    double* Fk = wsm + C::FRAG_D;                                     // series factors i a_i: reuse part[] (dead after the combine)
    double* Ck = Ok + C::P_D;                                         // series coefficients
    double* delta = Ck + C::C_D;                                      // delta[q * PD_EMAX + e]
    double* pre = delta + C::DELTA_D;                                 // prefactor of subset q
    unsigned long long* stride = reinterpret_cast<unsigned long long*>(pre + 4);   // mixed-radix stride of edge e
    int* verts = reinterpret_cast<int*>(stride + PD_EMAX);            // verts[e] = u_e, verts[E + e] = v_e
    int* reps = verts + 2 * PD_EMAX;
    const bool loop = (p.D != nullptr);
    constexpr int NK = C::NK, NT = C::NT;

    long long cur_pat = -1;
    int E = 0, T = 0, tp = 0, odd = -1, order = 0;
    unsigned long long steps = 0;
    const double2* Ap = p.A;
    const double2* Dp = p.D;

    for (;;) {
        unsigned long long chunk = 0;
        if (lane == 0) chunk = atomicAdd(p.counter, 1ull);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= p.nchunks) break;
        long long lo = 0, hi = p.B;                      // pattern owning this chunk: largest pat with coff[pat] <= chunk
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(p.coff + mid) <= chunk) lo = mid; else hi = mid;
        }
        const long long pat = lo;
        if (pat != cur_pat) {
            cur_pat = pat;
            const PatDesc* d = p.desc + pat;
            E = d->E; T = d->N / 2; steps = d->steps; odd = d->odd;
            order = odd >= 0 ? d->N : T;                     // series order: vertices for an odd total (f_loop_odd), pairs otherwise
            tp = TAIL ? E - 4 * TF : 0;
            Dp = (p.D && p.gidx) ? p.D + (size_t)__ldg(p.gidx + pat) * p.ldg : p.D;
            Ap = p.aidx ? p.A + (size_t)__ldg(p.aidx + pat) * p.lda * p.lda : p.A;
            __syncwarp();
            if (lane < PD_EMAX) {
                verts[lane] = lane < E ? d->u[lane] : 0;
                verts[PD_EMAX + lane] = lane < E ? d->v[lane] : 0;
                reps[lane] = lane < E ? d->r[lane] : 0;
            }
            if (lane == 0) {                             // stride_e = prod_{i > e} (r_i + 1)  (find_kept_edges, MSB first)
                unsigned long long s = 1;
                for (int e = E - 1; e >= 0; --e) { stride[e] = s; s *= (unsigned long long)d->r[e] + 1ull; }
            }
            __syncwarp();
            // fragment table of A'' (layout of haf_prep_kernel): element index a < E -> vertex u_a, a >= E -> v_(a-E)
            const bool packt = TAIL && tp == 1;
            for (int idx = lane; idx < NK * NT * 32; idx += 32) {
                const int pair = idx >> 5;
                const int tile = pair % NT, kappa = pair / NT;
                const int k = lane & 3, ncol = lane >> 2;
                const bool packk = packt && kappa == 2 * TF;
                const bool imag_slot = packk && k >= 2;
                const int ek = haf_chunk_elem(kappa, packk ? (k & 1) : k, E, TF, tp);
                double2 v = make_double2(0.0, 0.0);
                if (ek >= 0) {
                    const int vk = ek < E ? verts[ek] : verts[PD_EMAX + ek - E];
                    if (tile < TF) {
                        const int in = 4 * tile + (ncol >> 1);
                        if (in < E) {
                            const int vn = (ncol & 1) ? verts[PD_EMAX + in] : verts[in];
                            const double2 a = __ldg(Ap + (size_t)vk * p.lda + vn);
                            v = imag_slot ? make_double2(-a.y, a.x) : a;
                        }
                    } else {
                        const int tq = ncol >> 1;
                        if ((tq >> 1) < tp) {
                            const int in = 4 * TF + (tq >> 1);
                            const int vn = (tq & 1) ? verts[PD_EMAX + in] : verts[in];
                            const double2 a = __ldg(Ap + (size_t)vk * p.lda + vn);
                            v = (ncol & 1) ? make_double2(a.y, a.x) : make_double2(a.x, -a.y);
                            if (packk) v = make_double2(imag_slot ? v.y : v.x, 0.0);
                        }
                    }
                }
                sfrag[idx] = v;
            }
            __syncwarp();
        }
        const unsigned long long jb = (chunk - __ldg(p.coff + pat)) * BW_CHUNK;
        const unsigned long long je = jb + BW_CHUNK < steps ? jb + BW_CHUNK : steps;
        const int nprod = (T - 1) >> 1;                  // products: B_2 .. B_K, K = nprod + 1
        const int K = nprod + 1;                         // tr(M^j), j <= K, come from single elements
        const int nstepD = T >> 1;                       // products of the loop row (l_1 .. l_T)
        const bool packt = TAIL && tp == 1;
        cdd acc;
        acc.re = {0.0, 0.0};
        acc.im = {0.0, 0.0};

        for (unsigned long long jg = jb; jg < je; jg += 4) {
            // ---- decode the four subsets: lane (q8 = lane >> 3, e8 = lane & 7) handles edges e8 and e8 + 8 of subset q8
            {
                const int q8 = lane >> 3, e8 = lane & 7;
                const unsigned long long jj = jg + q8;
                int es = 0;
                double wt = 1.0;
                bool d0zero = false;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int e = e8 + 8 * rr;
                    double dv = 0.0;
                    if (e < E) {
                        const int r = reps[e];
                        const int kp = (int)((jj / stride[e]) % (unsigned long long)(r + 1));
                        es += kp;
                        wt *= binom_d(r, kp);
                        const int dlt = p.glynn ? 2 * kp - r : kp;
                        if (e == 0) d0zero = (dlt == 0);
                        dv = (double)dlt;
                    }
                    delta[q8 * PD_EMAX + e] = dv;
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    es += __shfl_xor_sync(0xffffffffu, es, off);
                    wt *= shfl_xor_d(wt, off);
                }
                const bool dz = __shfl_sync(0xffffffffu, d0zero ? 1 : 0, lane & ~7) != 0;
                if (e8 == 0) {
                    double pf = (((T - es) & 1) ? -1.0 : 1.0) * wt;      // (-1)^(N/2 - sum kept) prod C(r, kept)  (_hafnian.py:553-556)
                    if (p.glynn && dz && odd < 0) pf *= 0.5;                 // Glynn halving of edge 0: only without an unpaired vertex (:557-559)
                    pre[q8] = (jj < je) ? pf : 0.0;
                }
            }
            for (int s = lane; s < (T + 2) * 8; s += 32) part[s] = make_double2(0.0, 0.0);
            __syncwarp();
            double dl[TF > 0 ? TF : 1];
#pragma unroll
            for (int tau = 0; tau < TF; ++tau) {
                const int i = 4 * tau + t;
                dl[tau] = (i < E) ? delta[q * PD_EMAX + i] : 0.0;
            }
            double dlt = 0.0;
            if (TAIL && (t >> 1) < tp) dlt = delta[q * PD_EMAX + 4 * TF + (t >> 1)];

            const int npanels = E + (loop ? 1 : 0);
            for (int i = 0; i < npanels; ++i) {
                const bool isD = (i == E);
                const double rs = isD ? 1.0 : delta[q * PD_EMAX + i];
                if (!isD && __all_sync(0xffffffffu, rs == 0.0)) continue;     // edge deleted in all four subsets
                HafRow<TF, TAIL> w;
                HafY<TF, TAIL> y = {};
                const int v = i + half * E;                                   // row index in A'' (vertex pair i)
                // loop row: half 0 = D; half 1 = row `odd` of A (the unpaired vertex, f_loop_odd) or empty
                const bool oddrow = isD && half == 1 && odd >= 0;
                const int vv = isD ? (oddrow ? odd : 0) : (half ? verts[PD_EMAX + i] : verts[i]);
                {
                    const bool rowok = isD ? (half == 0 || oddrow) : true;
#pragma unroll
                    for (int tau = 0; tau < TF; ++tau) {
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const int iv = 4 * tau + t;
                            const bool ok = rowok && (iv < E);
                            double2 a = make_double2(0.0, 0.0);
                            if (ok) {
                                const int vc = r ? verts[PD_EMAX + iv] : verts[iv];
                                a = (isD && !oddrow) ? __ldg(Dp + vc) : __ldg(Ap + (size_t)vv * p.lda + vc);
                            }
                            w.wr[tau][r] = a.x;
                            w.wi[tau][r] = a.y;
                        }
                    }
                    w.wtr = w.wti = 0.0;
                    if (TAIL) {
                        const bool ok = rowok && ((t >> 1) < tp);
                        if (ok) {
                            const int iv = 4 * TF + (t >> 1);
                            const int vc = (t & 1) ? verts[PD_EMAX + iv] : verts[iv];
                            const double2 a = (isD && !oddrow) ? __ldg(Dp + vc) : __ldg(Ap + (size_t)vv * p.lda + vc);
                            w.wtr = a.x; w.wti = a.y;
                        }
                    }
                }
                (void)v;
                const bool in_tail = TAIL && (i >= 4 * TF);
                const int own_t = in_tail ? 2 * (i - 4 * TF) + (1 - half) : (i & 3);
                const int own_tau = i >> 2;
                double orr, oi, er, ei;
                if (!isD) {
                    if (K < 2) pat_advance<TF, TAIL, true, false>(w, y, dl, dlt, orr, oi, er, ei);
                    else pat_advance<TF, TAIL, false, false>(w, y, dl, dlt, orr, oi, er, ei);
                    if (t == 0) {
                        const int sv = half ? verts[i] : verts[PD_EMAX + i];                 // sigma(v)
                        const double2 a = __ldg(Ap + (size_t)vv * p.lda + sv);
                        double2 pp = part[1 * 8 + g];
                        pp.x += rs * a.x; pp.y += rs * a.y;
                        part[1 * 8 + g] = pp;
                        if (K < 2 && T >= 2) {
                            double2 p2 = part[2 * 8 + g];
                            p2.x += rs * er; p2.y += rs * ei;
                            part[2 * 8 + g] = p2;
                        }
                    }
                } else {
                    HafY<TF, TAIL> y0;  // l_1 = <Z_0, S Z_0>;  o_1 = <Z_0, S V_0> (V = chain of the odd vertex's row, half 1)
                    pat_advance<TF, TAIL, false, true>(w, y0, dl, dlt, orr, oi, er, ei);
                    y = y0;
                    pat_inner<TF, TAIL, true>(w, y, orr, oi);
                    if (t == 0 && half == 0) { Lk[(1 * 4 + q) * 2] = orr; Lk[(1 * 4 + q) * 2 + 1] = oi; }
                    if (odd >= 0) {
                        pat_inner<TF, TAIL, false>(w, y, er, ei);
                        if (t == 0 && half == 1) { Ok[(1 * 4 + q) * 2] = er; Ok[(1 * 4 + q) * 2 + 1] = ei; }
                    }
                }
                const int nsteps = isD ? nstepD : nprod;
                for (int k = 1; k <= nsteps; ++k) {
                    haf_step<TF, TAIL>(sfrag, lane, y, w, packt);   // w = row of B_(k+1) (Z_k for the loop row)
                    if (!isD) {
                        if (t == own_t) {                            // tr(M^(k+1)) share: element sigma(v) of this row
                            double dr, di;
                            if (in_tail) { dr = w.wtr; di = w.wti; }
                            else {
                                dr = 0.0; di = 0.0;
#pragma unroll
                                for (int tau = 0; tau < TF; ++tau)
                                    if (tau == own_tau) { dr = half ? w.wr[tau][0] : w.wr[tau][1]; di = half ? w.wi[tau][0] : w.wi[tau][1]; }
                            }
                            double2 pp = part[(k + 1) * 8 + g];
                            pp.x += rs * dr; pp.y += rs * di;
                            part[(k + 1) * 8 + g] = pp;
                        }
                        const bool needO = (2 * k + 1 > K) && (2 * k + 1 <= T);
                        const bool needE = (2 * k + 2 > K) && (2 * k + 2 <= T);
                        if (needO || needE) {
                            pat_advance<TF, TAIL, true, false>(w, y, dl, dlt, orr, oi, er, ei);
                            if (t == 0) {
                                if (needO) { double2 pp = part[(2 * k + 1) * 8 + g]; pp.x += rs * orr; pp.y += rs * oi; part[(2 * k + 1) * 8 + g] = pp; }
                                if (needE) { double2 pp = part[(2 * k + 2) * 8 + g]; pp.x += rs * er; pp.y += rs * ei; part[(2 * k + 2) * 8 + g] = pp; }
                            }
                        } else {
                            pat_advance<TF, TAIL, false, false>(w, y, dl, dlt, orr, oi, er, ei);
                        }
                    } else if (odd < 0) {  // l_(2k) = <Z_k, S Z_(k-1)>, l_(2k+1) = <Z_k, S Z_k>
                        pat_advance<TF, TAIL, true, true>(w, y, dl, dlt, orr, oi, er, ei);
                        if (t == 0 && half == 0) {
                            Lk[((2 * k) * 4 + q) * 2] = orr; Lk[((2 * k) * 4 + q) * 2 + 1] = oi;
                            Lk[((2 * k + 1) * 4 + q) * 2] = er; Lk[((2 * k + 1) * 4 + q) * 2 + 1] = ei;
                        }
                    } else {               // also o_(2k) = <Z_k, S V_(k-1)>, o_(2k+1) = <Z_k, S V_k>: the partner row's W against this row's Y
                        double cr, ci, c2r, c2i;
                        pat_inner<TF, TAIL, true>(w, y, orr, oi);
                        pat_inner<TF, TAIL, false>(w, y, cr, ci);
                        pat_advance<TF, TAIL, false, true>(w, y, dl, dlt, er, ei, c2r, c2i);      // y <- S w
                        pat_inner<TF, TAIL, true>(w, y, er, ei);
                        pat_inner<TF, TAIL, false>(w, y, c2r, c2i);
                        if (t == 0 && half == 0) {
                            Lk[((2 * k) * 4 + q) * 2] = orr; Lk[((2 * k) * 4 + q) * 2 + 1] = oi;
                            Lk[((2 * k + 1) * 4 + q) * 2] = er; Lk[((2 * k + 1) * 4 + q) * 2 + 1] = ei;
                        }
                        if (t == 0 && half == 1) {
                            Ok[((2 * k) * 4 + q) * 2] = cr; Ok[((2 * k) * 4 + q) * 2 + 1] = ci;
                            Ok[((2 * k + 1) * 4 + q) * 2] = c2r; Ok[((2 * k + 1) * 4 + q) * 2 + 1] = c2i;
                        }
                    }
                }
                __syncwarp();  // part[] slots are updated by different lanes in the next panel
            }
            // ---- combine the two rows (vertex e and its partner) of each subset q
            for (int s = lane; s < (T + 1) * 4; s += 32) {
                const int j = s >> 2, qq = s & 3;
                const double2 a = part[j * 8 + qq], b = part[j * 8 + qq + 4];
                Pk[(j * 4 + qq) * 2] = a.x + b.x; Pk[(j * 4 + qq) * 2 + 1] = a.y + b.y;
            }
            __syncwarp();
            // ---- series factors F_i = i a_i of exp(sum_i a_i eta^i).  Even total (f_loop, _hafnian.py:212-242): a_i = p_i/(2i) + l_i/2,
            // order T.  Odd total (f_loop_odd, :246-285): a_1 = oddloop, a_(2t) = p_t/(2t) + l_t/2, a_(2t+1) = o_t, order N = 2T + 1.
            {
                const double2 oddloop = odd >= 0 ? __ldg(Dp + odd) : make_double2(0.0, 0.0);
                for (int s = lane; s < order * 4; s += 32) {
                    const int i = (s >> 2) + 1, qq = s & 3;
                    double fr, fi;
                    if (odd < 0) {
                        fr = 0.5 * Pk[(i * 4 + qq) * 2]; fi = 0.5 * Pk[(i * 4 + qq) * 2 + 1];
                        if (loop) { fr += 0.5 * i * Lk[(i * 4 + qq) * 2]; fi += 0.5 * i * Lk[(i * 4 + qq) * 2 + 1]; }
                    } else if (i == 1) {
                        fr = oddloop.x; fi = oddloop.y;
                    } else if ((i & 1) == 0) {
                        const int tt = i >> 1;
                        fr = Pk[(tt * 4 + qq) * 2] + tt * Lk[(tt * 4 + qq) * 2]; fi = Pk[(tt * 4 + qq) * 2 + 1] + tt * Lk[(tt * 4 + qq) * 2 + 1];
                    } else {
                        const int tt = i >> 1;
                        fr = i * Ok[(tt * 4 + qq) * 2]; fi = i * Ok[(tt * 4 + qq) * 2 + 1];
                    }
                    Fk[(i * 4 + qq) * 2] = fr; Fk[(i * 4 + qq) * 2 + 1] = fi;
                }
                if (lane < 4) { Ck[lane * 2] = 1.0; Ck[lane * 2 + 1] = 0.0; }
                __syncwarp();
            }
            // ---- c_s = (1/s) sum_(i=1..s) F_i c_(s-i): the eight lanes of a subset (same q) split the sum over i and reduce by
            // shuffles (a serial version costs ~30 cycles per term on one lane: 1200 terms at order 49)
            {
                const int l8 = half * 4 + t;
                for (int sidx = 1; sidx <= order; ++sidx) {
                    double sr = 0.0, si = 0.0;
                    for (int i = 1 + l8; i <= sidx; i += 8) {
                        const double fr = Fk[(i * 4 + q) * 2], fi = Fk[(i * 4 + q) * 2 + 1];
                        const double c_r = Ck[((sidx - i) * 4 + q) * 2], c_i = Ck[((sidx - i) * 4 + q) * 2 + 1];
                        sr = fma(fr, c_r, sr); sr = fma(-fi, c_i, sr);
                        si = fma(fr, c_i, si); si = fma(fi, c_r, si);
                    }
                    sr += shfl_xor_d(sr, 1); si += shfl_xor_d(si, 1);
                    sr += shfl_xor_d(sr, 2); si += shfl_xor_d(si, 2);
                    sr += shfl_xor_d(sr, 16); si += shfl_xor_d(si, 16);
                    if (l8 == 0) { Ck[(sidx * 4 + q) * 2] = sr / sidx; Ck[(sidx * 4 + q) * 2 + 1] = si / sidx; }
                    __syncwarp();
                }
                if (l8 == 0) {
                    const double pf = pre[q];
                    if (pf != 0.0) {
                        dd_add(acc.re, pf * Ck[(order * 4 + q) * 2]);
                        dd_add(acc.im, pf * Ck[(order * 4 + q) * 2 + 1]);
                    }
                }
            }
            __syncwarp();
        }
        // fixed-order combine of the four per-subset accumulators (lanes 0, 4, 8, 12) -> one partial per chunk
        {
            cdd tot;
            tot.re = {0.0, 0.0};
            tot.im = {0.0, 0.0};
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                dd r, im;
                r.hi = shfl_d(acc.re.hi, 4 * qq); r.lo = shfl_d(acc.re.lo, 4 * qq);
                im.hi = shfl_d(acc.im.hi, 4 * qq); im.lo = shfl_d(acc.im.lo, 4 * qq);
                dd_add_dd(tot.re, r);
                dd_add_dd(tot.im, im);
            }
            if (lane == 0) {
                double* o = p.partial + chunk * 4;
                o[0] = tot.re.hi; o[1] = tot.re.lo; o[2] = tot.im.hi; o[3] = tot.im.lo;
            }
        }
    }
}

template <int TF, bool TAIL>
static int launch_pat_dmma(const PatParams& p, int sms, cudaStream_t st) {
    using C = PatCfg<TF, TAIL>;
    auto kern = pat_dmma_kernel<TF, TAIL>;
    WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
    unsigned long long want = (p.nchunks + C::WARPS - 1) / C::WARPS;
    const int grid = (int)(want < (unsigned long long)sms ? (want ? want : 1) : (unsigned long long)sms);
    kern<<<grid, 32 * C::WARPS, C::BYTES, st>>>(p);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}

static int launch_pat_class(int cls, const PatParams& p, int sms, cudaStream_t st) {
    switch (cls) {
        case 0: return launch_pat_dmma<1, false>(p, sms, st);
        case 1: return launch_pat_dmma<1, true>(p, sms, st);
        case 2: return launch_pat_dmma<2, false>(p, sms, st);
        case 3: return launch_pat_dmma<2, true>(p, sms, st);
        case 4: return launch_pat_dmma<3, false>(p, sms, st);
        case 5: return launch_pat_dmma<3, true>(p, sms, st);
        case 6: return launch_pat_dmma<4, false>(p, sms, st);
    }
    set_error("lhaf_patterns: bad tile class %d", cls);
    return WB200_EINVAL;
}

}  // namespace wb
