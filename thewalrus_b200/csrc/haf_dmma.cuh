// Register-resident row-panel machinery of the FP64 tensor-core (DMMA.8x8x4) hafnian kernels, shared by
// hafnian_dmma.cu (all edge repetitions 1: delta = +-1, sign masks) and pat_dmma.cuh (repetition patterns: integer
// delta incl. 0).  See hafnian_dmma.cu for the math and the register layout.
#pragma once
#include "common.cuh"

namespace wb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ double flipsign(double x, unsigned mask) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

// element held at K-chunk kappa, position k (or -1 for padding)
__device__ __forceinline__ int haf_chunk_elem(int kappa, int k, int m, int TF, int tp) {
    if (kappa < 2 * TF) {
        const int iv = 4 * (kappa >> 1) + k;
        return iv < m ? iv + (kappa & 1) * m : -1;
    }
    return (k >> 1) < tp ? 4 * TF + (k >> 1) + (k & 1) * m : -1;
}

// tile shape for m vertex pairs: TF full tiles of four pairs (+ a packed tail tile of one or two pairs)
static inline void haf_shape(int m, int* TF, int* tail) {
    const int r = m & 3;
    if (m >= 5 && (r == 1 || r == 2)) { *TF = m >> 2; *tail = 1; }
    else { *TF = (m + 3) >> 2; *tail = 0; }
}


// Fragment table, (kappa * NT + tile) * 32 + lane:
//   standard tile taup: (Ar, Ai)[e_k, e_n], e_n = 4 taup + (ncol >> 1) + (ncol & 1) m
//   tail tile        : (F1, F2) with column ncol <-> vertex u(ncol >> 1), output part ncol & 1:
//                      re: (Ar, -Ai), im: (Ai, Ar)   so that  D += yr * F1 + yi * F2
// Every CTA builds the table straight into its shared memory (it costs what copying a prebuilt table would, and
// saves the separate prep launch of round 1: 3 us + a launch gap on a 40 us n = 24 call).
__device__ __forceinline__ void haf_build_frag(const double* __restrict__ A, int n, int m, int TF, int tail,
                                               double2* __restrict__ frag, int tid, int nthreads) {
    const int tp = tail ? m - 4 * TF : 0;
    const int NK = 2 * TF + (tail ? 1 : 0), NT = TF + (tail ? 1 : 0);
    const int total = NK * NT * 32;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int lane = idx & 31, pair = idx >> 5;
        const int tile = pair % NT, kappa = pair / NT;
        const int k = lane & 3, ncol = lane >> 2;
        // K-packed tail chunk (one vertex pair in the tail, tp == 1): only positions 0, 1 of the chunk carry
        // elements, so positions 2, 3 take the IMAGINARY parts of the same two elements and the step needs two
        // DMMAs per tile for this chunk instead of four:  D_re += [yr | yi] . [Ar ; -Ai],  D_im += [yr | yi] . [Ai ; Ar]
        const bool packk = tail && tp == 1 && kappa == 2 * TF;
        const bool imag_slot = packk && k >= 2;
        const int ek = haf_chunk_elem(kappa, packk ? (k & 1) : k, m, TF, tp);
        double2 v = make_double2(0.0, 0.0);
        if (ek >= 0) {
            if (tile < TF) {
                const int in = 4 * tile + (ncol >> 1);
                if (in < m) {
                    const int en = in + (ncol & 1) * m;
                    const double ar = __ldg(A + 2 * ((size_t)ek * n + en)), ai = __ldg(A + 2 * ((size_t)ek * n + en) + 1);
                    v = imag_slot ? make_double2(-ai, ar) : make_double2(ar, ai);
                }
            } else {
                const int tq = ncol >> 1;
                if ((tq >> 1) < tp) {
                    const int en = 4 * TF + (tq >> 1) + (tq & 1) * m;
                    const double ar = __ldg(A + 2 * ((size_t)ek * n + en)), ai = __ldg(A + 2 * ((size_t)ek * n + en) + 1);
                    v = (ncol & 1) ? make_double2(ai, ar) : make_double2(ar, -ai);
                    if (packk) v = make_double2(imag_slot ? v.y : v.x, 0.0);   // D += [yr | yi] . [F1 ; F2]
                }
            }
        }
        frag[idx] = v;
    }
}

template <int TF, bool TAIL>
struct HafRow {  // one 8-row panel slice held by a thread
    double wr[TF > 0 ? TF : 1][2], wi[TF > 0 ? TF : 1][2];
    double wtr, wti;
};
template <int TF, bool TAIL>
struct HafY {
    double yr[TF > 0 ? 2 * TF : 1], yi[TF > 0 ? 2 * TF : 1];
    double ytr, yti;
};

// W <- Y * A' on the tensor pipe
template <int TF, bool TAIL>
__device__ __forceinline__ void haf_step(const double2* __restrict__ sfrag, int lane, const HafY<TF, TAIL>& y,
                                         HafRow<TF, TAIL>& w, bool packk) {
    constexpr int NK = 2 * TF + (TAIL ? 1 : 0), NT = TF + (TAIL ? 1 : 0);
#pragma unroll
    for (int tp = 0; tp < TF; ++tp) {
        w.wr[tp][0] = w.wr[tp][1] = 0.0;
        w.wi[tp][0] = w.wi[tp][1] = 0.0;
    }
    w.wtr = w.wti = 0.0;
#pragma unroll
    for (int kap = 0; kap < NK; ++kap) {
        const double ar = kap < 2 * TF ? y.yr[kap < 2 * TF ? kap : 0] : y.ytr;
        const double ai = kap < 2 * TF ? y.yi[kap < 2 * TF ? kap : 0] : y.yti;
        if (TAIL && kap == 2 * TF && packk) {   // K-packed tail chunk (see haf_prep_kernel): 2 DMMAs per tile
            const double yi2 = __shfl_sync(0xffffffffu, y.yti, lane & ~2);
            const double ap = (lane & 2) ? yi2 : y.ytr;
#pragma unroll
            for (int tp = 0; tp < TF; ++tp) {
                const double2 b = sfrag[(kap * NT + tp) * 32 + lane];
                dmma884(w.wr[tp][0], w.wr[tp][1], ap, b.x);
                dmma884(w.wi[tp][0], w.wi[tp][1], ap, b.y);
            }
            const double2 b = sfrag[(kap * NT + TF) * 32 + lane];
            dmma884(w.wtr, w.wti, ap, b.x);
            continue;
        }
#pragma unroll
        for (int tp = 0; tp < TF; ++tp) {
            const double2 b = sfrag[(kap * NT + tp) * 32 + lane];
            const double nbi = -b.y;
            dmma884(w.wr[tp][0], w.wr[tp][1], ar, b.x);
            dmma884(w.wi[tp][0], w.wi[tp][1], ar, b.y);
            dmma884(w.wr[tp][0], w.wr[tp][1], ai, nbi);
            dmma884(w.wi[tp][0], w.wi[tp][1], ai, b.x);
        }
        if (TAIL) {
            const double2 b = sfrag[(kap * NT + TF) * 32 + lane];
            dmma884(w.wtr, w.wti, ar, b.x);
            dmma884(w.wtr, w.wti, ai, b.y);
        }
    }
}

#define WB_CFMA(sr, si, xr_, xi_, yr_, yi_)                  \
    do {                                                       \
        sr = fma(xr_, yr_, sr); sr = fma(-(xi_), yi_, sr);     \
        si = fma(xr_, yi_, si); si = fma(xi_, yr_, si);        \
    } while (0)

// Y <- S_j W (swap partners, sign delta).  If IP: also odd = <X, Y_old>, even = <X, Y_new> over this
// thread's slots, X = W of the partner row (lane ^ 16) or W itself (SELF, the loop row).
template <int TF, bool TAIL, bool IP, bool SELF>
__device__ __forceinline__ void haf_advance(const HafRow<TF, TAIL>& w, HafY<TF, TAIL>& y, const unsigned* sm,
                                            unsigned smt, double& orr, double& oi, double& er, double& ei) {
    double o2r = 0.0, o2i = 0.0, e2r = 0.0, e2i = 0.0;  // second chains for ILP
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
        y.yr[2 * tau + 0] = flipsign(w.wr[tau][1], sm[tau]);
        y.yr[2 * tau + 1] = flipsign(w.wr[tau][0], sm[tau]);
        y.yi[2 * tau + 0] = flipsign(w.wi[tau][1], sm[tau]);
        y.yi[2 * tau + 1] = flipsign(w.wi[tau][0], sm[tau]);
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
        y.ytr = flipsign(shfl_xor_d(w.wtr, 1), smt);
        y.yti = flipsign(shfl_xor_d(w.wti, 1), smt);
        if (IP) WB_CFMA(er, ei, xr, xi, y.ytr, y.yti);
    }
    if (IP) {
        orr += o2r; oi += o2i; er += e2r; ei += e2i;
        orr += shfl_xor_d(orr, 1); oi += shfl_xor_d(oi, 1); er += shfl_xor_d(er, 1); ei += shfl_xor_d(ei, 1);
        orr += shfl_xor_d(orr, 2); oi += shfl_xor_d(oi, 2); er += shfl_xor_d(er, 2); ei += shfl_xor_d(ei, 2);
    }
}


}  // namespace wb
