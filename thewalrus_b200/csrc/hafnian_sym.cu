// Hafnian, all edge repetitions 1, Glynn sieve — SYMMETRIC-HALF tensor-core kernel for even n from 36 to 64.
//
// Same sum as haf_dmma_kernel (thewalrus/_hafnian.py:416-467 + charpoly.powertrace + f), same row-panel products
// row c of B_(k+1) = (row c of B_k) S_j A'  on DMMA.8x8x4 with trace pairing.  What is new: B_k = M_j^(k-1) A' is
// SYMMETRIC, so of every product only the tiles on and above the diagonal are computed:
// the row panel of vertex pair i (tile tau_i = i / 4) computes the N-tiles tau' >= tau_i (+ the packed tail tile) and
// the tiles below the diagonal are the transposes of what the other panels computed.  At n = 50 that is 96.5 tile
// units per product and group of four subsets instead of 162.5 (0.59 of the DMMAs).
//
// That couples the panels of a group: all of them must finish product k before any starts product k + 1, and every
// panel needs entries other panels computed.  So the iterates of the groups in flight live in shared memory
// (four subsets x n x (n + 1) complex = 164 KB at n = 50, next to the 46 KB fragment table of A') and the warps of a team
// advance their group together, two barriers per product: [all panels read their rows of B_k and compute] | barrier |
// [write the computed tiles] | barrier.  A panel writes ONLY what it computed (its rows, the columns of tiles >= its own
// and the tail columns); the entries of earlier tiles are read column-wise from the rows of the panels that computed
// them (row stride n + 1, subset stride = 16 words mod 32: both access patterns are bank-conflict free).  The first
// barrier is split-phase (mbarrier): arrive after the K loop, wait before the stores, the trace pairing in between.
//
// Work split: each warp owns two row panels whose tile counts add up to the same number for every warp (HsShape: roles)
// and walks K ONCE for both - the later tile's panel needs a subset of the other's fragments, so every 16-byte fragment
// load feeds both and two independent panels' DMMAs interleave.  The K loop is rolled (one iteration = one K tile, the
// next tile's row entries prefetched): fully unrolled, ptxas hoists ~30 fragment loads and spills the accumulators.  The
// tail panel (vertex pair 24 at n = 50: only its 2 x 2 tail block is new) is split over K across the team's warps and
// summed by the team's first warp while the others store.
//
// The power traces stay warp-local although no warp ever sees a whole row of B_(k+1): with U = B_a S, V = B_b,
//     tr(M^(a+b)) = sum_(x,y) V[x][y] delta_x U[sigma(x)][y]
// and the term of (y, x) equals the term of (x, y) (both B_a and B_b are symmetric), so every COMPUTED entry (x, y) of
// a strictly-upper tile counts twice and the entries of the diagonal tiles once - the same partner-row inner products
// as haf_advance, restricted to the computed tiles, with a weight.  (NumPy emulation of exactly this bookkeeping
// against the oracle: 3e-15)  Per product every warp leaves its trace shares in shared memory; after
// the barrier warp w sums the shares of one (trace kind, subset) in a fixed shuffle tree.  The series of a group
// (thewalrus/_hafnian.py:183-214, f) runs on the team's first warp in "push" form, one independent chain per subset.
//
// Shapes (HsShape, haf_sym_launch): whole tiles of four vertex pairs plus at most one packed tail pair - n = 40, 42, 48, 50
// (two teams of 6 warps, two subsets each), 56, 58, 64 (one team of 8 warps, two subsets); the sizes in between run
// zero-padded in the next whole-tile shape.  Loop hafnians, other sizes and short ranges stay on haf_dmma_kernel.
//
// Measured on B200 (profiles/r02_haf_sym_vs_panel.txt, r02_ncu_haf50_sym.txt): n = 50: 3.87e6 subsets/s against 2.75e6 of
// the row-panel kernel (1.41x) at 0.59 of its DMMAs, n = 56: 2.86e6 against 1.90e6 (1.51x) - tensor pipe 78 - 81 % busy
// (row-panel kernel: 92 %): the barriers, the store phase and the pairing phase of all warps at once cost what the
// row-panel kernel overlaps (profiles/r02_haf_sym_phase_costs.txt).
#include <stdlib.h>
#include "common.cuh"
#include "haf_dmma.cuh"

namespace wb {

// Shape of an instance: TF full tiles of four vertex pairs (+ a one-pair packed tail), NQ subsets per group.
//
// Teams.  NQ = 4: the warps advance one group of four subsets (a DMMA's 8 rows = 4 subsets x the two vertices of a pair).
// NQ = 2: 8 rows = 2 subsets x two vertex pairs x 2 - half the shared memory per group, so either two teams per CTA, each on
// its own group (TF = 6, the default for n = 48 / 50: every barrier spans 6 warps instead of 12), or a size that does not
// fit with four subsets at all (TF = 7, n = 56: one team).
// The hope for two teams was more: teams running out of step, one team's barriers / store / pairing phases under the
// other's DMMAs.  Measured (profiles/r02_haf_sym_skew.txt): holding the second team back by any fraction of a product changes
// n = 50 by +1 % and n = 48 by -5 % - six warps cannot be spread evenly over four schedulers (2 + 1 + 2 + 1), so while one
// team is in a phase the other team's warps on the 2-warp schedulers set its pace and the 1-warp schedulers idle.
//
// Roles.  A warp owns the panels of two tiles whose computed-tile counts add up to the same number for every warp:
// TF even: tiles r and TF - 1 - r (TF + 1 tile units); TF odd: tile 0 alone, then tiles r and TF - r (TF tile units).
template <int TF_, bool TAIL_, int NQ_, int TEAMS_>
struct HsShape {
    static constexpr int TF = TF_, NQ = NQ_;
    static constexpr bool TAIL = TAIL_;
    static constexpr int NK = 2 * TF + (TAIL ? 1 : 0);
    static constexpr int NT = TF + (TAIL ? 1 : 0);
    static constexpr int M = 4 * TF + (TAIL ? 1 : 0);              // vertex pairs
    static constexpr int N = 2 * M;
    static constexpr int PP = 4 / NQ;                              // vertex pairs per row panel
    static constexpr int SUBS = NQ;                                // panels per tile
    static constexpr int ROLES = (TF + 1) / 2;
    static constexpr int TW = ROLES * SUBS;                        // warps per team
    static constexpr int TEAMS = TEAMS_;                           // groups in flight per CTA (shared memory decides)
    static constexpr int WARPS = TEAMS * TW;
    static constexpr int FRAG_D = NK * NT * 64;                    // doubles
    static constexpr int LD = N + 1;                               // row stride of the state (entries): 4 LD = 4 or 12 (mod 32) words
    static constexpr int QS = ((N * LD + 7) / 8) * 8 + 4;          // subset stride: 4 QS = 16 (mod 32) words -> the two subsets of a
                                                                   // quarter warp never share a bank, row-wise or column-wise
    static constexpr int STATE_D2 = TEAMS * NQ * QS;               // double2: state[team][subset][row][col], only tiles >= the row's tile valid
    static constexpr int P_D = TEAMS * NQ * (M + 2) * 2;           // P[team][subset][j] complex (doubles)
    static constexpr int TMP_D = TEAMS * 3 * TW * 8 * 2;           // ptmp[team][kind][warp in team][row] complex (doubles)
    static constexpr int TAILC_D = TAIL ? WARPS * 32 * 2 : 0;      // partial tail tiles of the K-split tail panel (doubles)
    static constexpr size_t BYTES = sizeof(double) * ((size_t)FRAG_D + 2 * (size_t)STATE_D2 + P_D + TMP_D + TAILC_D);
    __host__ __device__ static constexpr int roleA(int r) { return r; }
    __host__ __device__ static constexpr int roleB(int r) { return (TF & 1) ? (r == 0 ? TF : TF - r) : TF - 1 - r; }      // TF: no second panel
    __device__ static __forceinline__ int team(int warp) { return TEAMS == 1 ? 0 : (((warp >> 2) + warp) & 1); }
    __device__ static __forceinline__ int wl(int warp) { return TEAMS == 1 ? warp : (warp >> 1); }
    __device__ static __forceinline__ int q(int lane) { return (lane >> 2) & (NQ - 1); }
    __device__ static __forceinline__ int ps(int lane) { return NQ == 4 ? 0 : ((lane >> 3) & 1); }
    __device__ static __forceinline__ void sync(int team) {
        if (TEAMS == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" :: "r"(1 + team), "r"(32 * TW) : "memory");
    }
};

// sign mask of vertex pair p: neg has bit p set where delta_p = -1 (pairs beyond the true size: delta = +1)
template <class S>
__device__ __forceinline__ unsigned hs_sign(unsigned neg, int p) {
    return ((neg >> p) & 1u) << 31;
}
// Sizes between the shapes: a matrix of mt < S::M vertex pairs runs in the next shape with whole tiles as if padded with
// zero rows and columns (they change no trace); only the first product, which reads A' itself, the signs and the series
// know the true size.
__device__ __forceinline__ unsigned hs_negmask(uint64_t jq, int mt) {
    const unsigned kept = __brev((unsigned)jq) >> (32 - mt);         // bit p = bit mt - 1 - p of jq (thewalrus/_hafnian.py:162-180)
    return ~kept & (mt >= 32 ? 0xffffffffu : (1u << mt) - 1u);          // mt <= 32: jq < 2^31
}

// One row of Y_k = B_k S, read just in time: chunk kap, position t of the lane.  Y_k[v][c] = delta_c B_k[v][sigma(c)], with
// B_1 = A' (global memory, k = 1) or B_k in the shared-memory state.  Holding the row in registers (as haf_dmma_kernel
// does) on top of the accumulators of two panels does not fit 168 registers.
template <class S, int FIRST>
struct HsY {
    const double2* row;      // row of this lane's (subset, vertex): state (stride LD) or A' (stride n, k = 1)
    const double2* col;      // the same vertex as a COLUMN of the state: entry [c][v] = col[c * LD]
    unsigned neg;            // bit p: delta of vertex pair p is -1
    int t;
    int mt;                  // FIRST only: true number of vertex pairs (partner offset and row length in A')
    bool rowok;              // FIRST only: this lane's row is a real vertex (not padding)
    static constexpr bool first = FIRST != 0;   // k = 1 (FIRST = 1; 2: of a zero-padded size): B_1 = A' (global memory), every entry is there, read row-wise;
                                           // a compile-time property, so that the loads of k > 1 are plain LDS
    __device__ __forceinline__ void get(int kap, double& yr, double& yi) const {   // computed tiles only (row-wise)
        constexpr int TF = S::TF;
        const int m = FIRST ? mt : S::M;
        if (kap < 2 * TF) {
            const int tau = kap >> 1;
            const bool ok = FIRST != 2 || (rowok && 4 * tau + t < mt);
            const double2 a = ok ? row[4 * tau + t + (1 - (kap & 1)) * m] : make_double2(0.0, 0.0);
            const unsigned s = hs_sign<S>(neg, 4 * tau + t);
            yr = flipsign(a.x, s); yi = flipsign(a.y, s);
        } else {
            yr = yi = 0.0;
            if (S::TAIL && t < 2) {
                const double2 a = row[4 * TF + (1 - (t & 1)) * m];
                const unsigned s = hs_sign<S>(neg, 4 * TF);
                yr = flipsign(a.x, s); yi = flipsign(a.y, s);
            }
        }
    }
};

// Walks the row of B_k of one lane K tile by K tile (a0 = entry of column 4 tau + t + m, a1 = of column 4 tau + t).
// A panel of tile T computed and stored only the columns of tiles >= T of its rows; the entries of earlier tiles were
// computed by other panels as THEIR rows, and B_k is symmetric: read them column-wise (conflict-free with the padded
// strides).  Nothing is stored twice (the first version stored every strictly-upper tile also transposed: 6 % of the kernel).
template <class S, int FIRST, int T>
struct HsCur {
    const double2* p;
    int d, inc;
    __device__ __forceinline__ void init(const HsY<S, FIRST>& y) {
        constexpr int m = S::M, LD = S::LD;
        if (FIRST) { p = y.row + y.t; d = y.mt; inc = 4; }
        else if (T == 0) { p = y.row + y.t; d = m; inc = 4; }
        else { p = y.col + y.t * LD; d = m * LD; inc = 4 * LD; }
    }
    // entries of K tile tau: a0 = column 4 tau + t + m (the partner half), a1 = column 4 tau + t
    __device__ __forceinline__ double2 a0(const HsY<S, FIRST>& y, int tau) const {
        return (FIRST != 2 || (y.rowok && 4 * tau + y.t < y.mt)) ? p[d] : make_double2(0.0, 0.0);
    }
    __device__ __forceinline__ double2 a1(const HsY<S, FIRST>& y, int tau) const {
        return (FIRST != 2 || (y.rowok && 4 * tau + y.t < y.mt)) ? p[0] : make_double2(0.0, 0.0);
    }
    __device__ __forceinline__ void next(const HsY<S, FIRST>& y, int tau_next) {
        constexpr int m = S::M;
        if (T > 0 && T < S::TF && !y.first && tau_next == T) { p = y.row + 4 * T + y.t; d = m; inc = 4; }
        else p += inc;
    }
    __device__ __forceinline__ double2 tail(const HsY<S, FIRST>& y) const {    // tail columns: every panel computes them itself
        return y.row[4 * S::TF + (1 - (y.t & 1)) * S::M];
    }
};

// Split-phase barrier (mbarrier in shared memory, one arrival per warp): "every panel has read its rows of B_k" is true
// when every warp has LEFT its K loop, but a warp needs it only when it is about to store - after its trace pairing,
// which reads nothing but the warp's own rows.  Arriving before the pairing and waiting after it hides the skew between
// the warps' K loops under the pairing phase.
__device__ __forceinline__ void hs_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hs_mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.release.cta.shared::cta.b64 st, [%0]; }" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void hs_mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// The tail panel (the one vertex pair beyond the full tiles, n = 50): only its 2 x 2 tail block has to be computed, every
// other entry of its rows arrives by symmetry.  That is 25 DMMAs in one dependent chain - on one warp it made that warp
// late at every barrier - so the K range is SPLIT over the warps of the team: warp w multiplies the K chunks w, w + TW, ...
// (the first warp also the packed tail chunk), the partial tiles meet in shared memory and the first warp sums them in
// warp order while the others store.
template <class S, int FIRST>
__device__ __forceinline__ double2 hs_tail_chunk(const double2* __restrict__ sfrag, int lane, int wl, const HsY<S, FIRST>& y) {
    constexpr int TF = S::TF, NT = S::NT, m = S::M, LD = S::LD, CH = (2 * TF + S::TW - 1) / S::TW;
    double pr = 0.0, pi = 0.0, p2r = 0.0, p2i = 0.0;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
        const int kap = wl + cc * S::TW, tau = kap >> 1, h = kap & 1;
        if (kap < 2 * TF) {
            const int c = 4 * tau + y.t + (1 - h) * m;                  // Y[v][chunk position] = delta B[v][c]
            const double2 a = y.first ? y.row[c] : y.col[c * LD];
            const unsigned s = hs_sign<S>(y.neg, 4 * tau + y.t);
            const double2 b = sfrag[(kap * NT + TF) * 32 + lane];
            dmma884(pr, pi, flipsign(a.x, s), b.x);
            dmma884(p2r, p2i, flipsign(a.y, s), b.y);
        }
    }
    if (wl == 0) {
        const double2 at = y.row[4 * TF + (1 - (y.t & 1)) * m];
        const unsigned st = (y.t < 2) ? hs_sign<S>(y.neg, 4 * TF) : 0u;
        const double ar = (y.t < 2) ? flipsign(at.x, st) : 0.0, ai = (y.t < 2) ? flipsign(at.y, st) : 0.0;
        const double yi2 = __shfl_sync(0xffffffffu, ai, lane & ~2);
        const double ap = (lane & 2) ? yi2 : ar;
        const double2 bt = sfrag[(2 * TF * NT + TF) * 32 + lane];
        dmma884(pr, pi, ap, bt.x);
    }
    return make_double2(pr + p2r, pi + p2i);
}

// The two panels of a warp (tiles TA and TB > TA; TB == TF: only one panel) in ONE pass over K: the panel of the later tile
// needs a subset of the fragments of the other, so every 16-byte fragment load feeds both, and the DMMAs of two independent
// panels interleave (a one-tile panel alone is a chain of dependent DMMAs).
// The loop over the K tiles is deliberately NOT unrolled: fully unrolled, ptxas hoists some thirty 16-byte fragment
// loads to the top of the block (130 registers) and spills the accumulators.
template <class S, int FIRST, int TA, int TB>
__device__ __forceinline__ void hs_step2(const double2* __restrict__ sfrag, int lane, const HsY<S, FIRST>& yA, const HsY<S, FIRST>& yB,
                                         HafRow<S::TF, S::TAIL>& wA, HafRow<S::TF, S::TAIL>& wB) {
    constexpr int TF = S::TF, NT = S::NT;
    constexpr bool TAIL = S::TAIL, HASB = TB < TF;
#pragma unroll
    for (int tp = TA; tp < TF; ++tp) {
        wA.wr[tp][0] = wA.wr[tp][1] = 0.0;
        wA.wi[tp][0] = wA.wi[tp][1] = 0.0;
    }
#pragma unroll
    for (int tp = TB; tp < TF; ++tp) {
        wB.wr[tp][0] = wB.wr[tp][1] = 0.0;
        wB.wi[tp][0] = wB.wi[tp][1] = 0.0;
    }
    wA.wtr = wA.wti = 0.0;
    wB.wtr = wB.wti = 0.0;
    const double2* fr = sfrag + lane;
    HsCur<S, FIRST, TA> curA;
    HsCur<S, FIRST, HASB ? TB : TA> curB;
    curA.init(yA); curB.init(yB);
    unsigned nb = yA.neg >> yA.t;                // bit 0: delta of the lane's column pair in the current K tile
    double2 a0 = curA.a0(yA, 0), a1 = curA.a1(yA, 0), c0 = make_double2(0.0, 0.0), c1 = c0;     // fetched one iteration ahead
    if (HASB) { c0 = curB.a0(yB, 0); c1 = curB.a1(yB, 0); }
#pragma unroll 1
    for (int tau = 0; tau < TF; ++tau) {
        const unsigned s = (nb & 1u) << 31;
        // next K tile; after the last one: the tail element of lanes t = 0, 1 (an address every lane may read)
        const bool last = tau == TF - 1;
        curA.next(yA, tau + 1);
        if (HASB) curB.next(yB, tau + 1);
        double2 n0 = make_double2(0.0, 0.0), n1 = n0, d0 = n0, d1 = n0;
        if (!last) {
            n0 = curA.a0(yA, tau + 1); n1 = curA.a1(yA, tau + 1);
            if (HASB) { d0 = curB.a0(yB, tau + 1); d1 = curB.a1(yB, tau + 1); }
        } else if (TAIL) {
            n0 = curA.tail(yA);
            if (HASB) d0 = curB.tail(yB);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double ar = flipsign(h ? a1.x : a0.x, s), ai = flipsign(h ? a1.y : a0.y, s);
            const double cr = flipsign(h ? c1.x : c0.x, s), ci = flipsign(h ? c1.y : c0.y, s);
#pragma unroll
            for (int tp = TA; tp < TF; ++tp) {
                const double2 b = fr[tp * 32];
                const double nbi = -b.y;
                dmma884(wA.wr[tp][0], wA.wr[tp][1], ar, b.x);
                dmma884(wA.wi[tp][0], wA.wi[tp][1], ar, b.y);
                if (tp >= TB) {
                    dmma884(wB.wr[tp][0], wB.wr[tp][1], cr, b.x);
                    dmma884(wB.wi[tp][0], wB.wi[tp][1], cr, b.y);
                }
                dmma884(wA.wr[tp][0], wA.wr[tp][1], ai, nbi);
                dmma884(wA.wi[tp][0], wA.wi[tp][1], ai, b.x);
                if (tp >= TB) {
                    dmma884(wB.wr[tp][0], wB.wr[tp][1], ci, nbi);
                    dmma884(wB.wi[tp][0], wB.wi[tp][1], ci, b.x);
                }
            }
            if (TAIL) {
                const double2 b = fr[TF * 32];
                dmma884(wA.wtr, wA.wti, ar, b.x);
                if (HASB) dmma884(wB.wtr, wB.wti, cr, b.x);
                dmma884(wA.wtr, wA.wti, ai, b.y);
                if (HASB) dmma884(wB.wtr, wB.wti, ci, b.y);
            }
            fr += NT * 32;
        }
        a0 = n0; a1 = n1; c0 = d0; c1 = d1;
        nb >>= 4;
    }
    if (TAIL) {                                  // K-packed tail chunk (one vertex pair in the tail): 2 DMMAs per tile
        const bool own = yA.t < 2;
        const unsigned s = own ? hs_sign<S>(yA.neg, 4 * TF) : 0u;
        const double ar = own ? flipsign(a0.x, s) : 0.0, ai = own ? flipsign(a0.y, s) : 0.0;
        const double cr = own ? flipsign(c0.x, s) : 0.0, ci = own ? flipsign(c0.y, s) : 0.0;
        const double ai2 = __shfl_sync(0xffffffffu, ai, lane & ~2), ci2 = __shfl_sync(0xffffffffu, ci, lane & ~2);
        const double ap = (lane & 2) ? ai2 : ar, cp = (lane & 2) ? ci2 : cr;
#pragma unroll
        for (int tp = TA; tp < TF; ++tp) {
            const double2 b = fr[tp * 32];
            dmma884(wA.wr[tp][0], wA.wr[tp][1], ap, b.x);
            dmma884(wA.wi[tp][0], wA.wi[tp][1], ap, b.y);
            if (tp >= TB) {
                dmma884(wB.wr[tp][0], wB.wr[tp][1], cp, b.x);
                dmma884(wB.wi[tp][0], wB.wi[tp][1], cp, b.y);
            }
        }
        const double2 b = fr[TF * 32];
        dmma884(wA.wtr, wA.wti, ap, b.x);
        if (HASB) dmma884(wB.wtr, wB.wti, cp, b.x);
    }
}

// Local pairing sums of one panel over its computed tiles: odd = <W(partner row), Y_old(own row)>, even = the same with
// Y_new = S W(own row); strictly-upper tiles weigh 2, the diagonal tile 1 (see the header).  Not reduced over lanes.
template <class S, int FIRST, int TAU>
__device__ __forceinline__ void hs_pairing(const HafRow<S::TF, S::TAIL>& w, const HsY<S, FIRST>& y,
                                           double& orr, double& oi, double& er, double& ei) {
    constexpr int TF = S::TF;
    orr = oi = er = ei = 0.0;
#pragma unroll
    for (int tau = TAU; tau < TF; ++tau) {
        const double wgt = tau == TAU ? 1.0 : 2.0;        // strictly-upper tiles count twice (their transposes are never computed)
        const double x0r = wgt * shfl_xor_d(w.wr[tau][0], 16), x0i = wgt * shfl_xor_d(w.wi[tau][0], 16);
        const double x1r = wgt * shfl_xor_d(w.wr[tau][1], 16), x1i = wgt * shfl_xor_d(w.wi[tau][1], 16);
        const unsigned s = hs_sign<S>(y.neg, 4 * tau + y.t);
        double yr, yi;
        y.get(2 * tau, yr, yi);
        WB_CFMA(orr, oi, x0r, x0i, yr, yi);
        y.get(2 * tau + 1, yr, yi);
        WB_CFMA(orr, oi, x1r, x1i, yr, yi);
        WB_CFMA(er, ei, x0r, x0i, flipsign(w.wr[tau][1], s), flipsign(w.wi[tau][1], s));   // Y_new, chunk 2 tau
        WB_CFMA(er, ei, x1r, x1i, flipsign(w.wr[tau][0], s), flipsign(w.wi[tau][0], s));   // Y_new, chunk 2 tau + 1
    }
    if (S::TAIL) {
        const double wgt = TAU == TF ? 1.0 : 2.0;
        const double xr = wgt * shfl_xor_d(w.wtr, 16), xi = wgt * shfl_xor_d(w.wti, 16);
        const unsigned smt = (y.t < 2) ? hs_sign<S>(y.neg, 4 * TF) : 0u;
        const double nr = flipsign(shfl_xor_d(w.wtr, 1), smt), ni = flipsign(shfl_xor_d(w.wti, 1), smt);
        double ytr, yti;
        y.get(2 * TF, ytr, yti);
        WB_CFMA(orr, oi, xr, xi, ytr, yti);
        WB_CFMA(er, ei, xr, xi, nr, ni);
    }
}

// the rows of Y_k = B_k S of a panel (first vertex pair ibase) as seen by this lane: from A' at k = 1 (B_1 = A', global
// memory), else from the team's shared-memory state
template <class S, int FIRST>
__device__ __forceinline__ HsY<S, FIRST> hs_rows(const double2* __restrict__ state, const double* __restrict__ A, int ibase, bool tailrows,
                                                 unsigned neg, int mt, int lane) {
    constexpr int m = S::M;
    const int half = lane >> 4;
    const int i = ibase + (tailrows ? 0 : S::ps(lane));                 // tail panel: the rows of the second pair slot repeat the first
    const int v = i + half * m;
    HsY<S, FIRST> y;
    y.mt = FIRST == 2 ? mt : m; y.rowok = FIRST != 2 || i < mt;
    y.row = FIRST ? reinterpret_cast<const double2*>(A) + (y.rowok ? (size_t)(i + half * y.mt) * (2 * y.mt) : 0)
                  : state + S::q(lane) * S::QS + v * S::LD;
    y.col = state + S::q(lane) * S::QS + v;
    y.neg = neg; y.t = lane & 3;
    return y;
}

// trace shares of one computed panel (first vertex pair ibase, tile TAU), added to the per-lane sums
template <class S, int FIRST, int TAU>
__device__ __forceinline__ void hs_traces(const HafRow<S::TF, S::TAIL>& w, const HsY<S, FIRST>& y, int ibase, bool needO, bool needE, int lane,
                                          double (&tr)[6]) {
    constexpr int TF = S::TF;
    constexpr bool in_tail = S::TAIL && TAU == TF;
    const int t = lane & 3, half = lane >> 4;
    const int i = ibase + (in_tail ? 0 : S::ps(lane));
    double rs = ((y.neg >> i) & 1u) ? -1.0 : 1.0;                       // delta of the row's vertex pair
    if (in_tail && S::ps(lane)) rs = 0.0;                               // repeated rows of the tail panel
    {   // tr(M^(k+1)) share: element sigma(v) of this row, in the diagonal tile
        const int own_t = in_tail ? (1 - half) : (i & 3);
        if (t == own_t) {
            double dr, di;
            if (in_tail) { dr = w.wtr; di = w.wti; }
            else { dr = half ? w.wr[TAU < TF ? TAU : 0][0] : w.wr[TAU < TF ? TAU : 0][1]; di = half ? w.wi[TAU < TF ? TAU : 0][0] : w.wi[TAU < TF ? TAU : 0][1]; }
            tr[0] += rs * dr; tr[1] += rs * di;
        }
    }
    if (needO || needE) {
        double orr, oi, er, ei;
        hs_pairing<S, FIRST, TAU>(w, y, orr, oi, er, ei);
        tr[2] += rs * orr; tr[3] += rs * oi;
        tr[4] += rs * er; tr[5] += rs * ei;
    }
}

// write the computed tiles of a panel (its rows, the columns of tiles >= its own and the tail columns)
template <class S, int TAU>
__device__ __forceinline__ void hs_store(double2* __restrict__ state, int ibase, int lane, const HafRow<S::TF, S::TAIL>& w) {
    constexpr int TF = S::TF, m = S::M;
    const int t = lane & 3, half = lane >> 4;
    if (TAU == TF && S::ps(lane)) return;                               // repeated rows of the tail panel
    double2* row = state + S::q(lane) * S::QS + (ibase + (TAU == TF ? 0 : S::ps(lane)) + half * m) * S::LD;
#pragma unroll
    for (int tau = TAU; tau < TF; ++tau) {
#pragma unroll
        for (int r = 0; r < 2; ++r) row[4 * tau + t + r * m] = make_double2(w.wr[tau][r], w.wi[tau][r]);
    }
    if (S::TAIL && t < 2) row[4 * TF + (t & 1) * m] = make_double2(w.wtr, w.wti);
}

// one product step of a warp with role RHO: its two panels (+ its K chunks of the tail panel)
template <class S, int FIRST, int RHO>
__device__ __forceinline__ void hs_warp_step(const double2* __restrict__ sfrag, double2* __restrict__ state, const double* __restrict__ A,
                                             int sub, bool needO, bool needE, bool store, unsigned neg, int mt,
                                             int lane, int team, int wl, double2* __restrict__ tailC, uint64_t* barA, unsigned parity, double (&tr)[6]) {
    constexpr int TF = S::TF, TA = S::roleA(RHO), TB = S::roleB(RHO);
    constexpr bool TAIL = S::TAIL, HASB = TB < TF;
    const int iA = 4 * TA + S::PP * sub, iB = HASB ? 4 * TB + S::PP * sub : iA;
    HafRow<TF, TAIL> wA, wB;
    HsY<S, FIRST> yC;
    if (TAIL) {
        yC = hs_rows<S, FIRST>(state, A, 4 * TF, true, neg, mt, lane);
        tailC[wl * 32 + lane] = hs_tail_chunk<S, FIRST>(sfrag, lane, wl, yC);
    }
    const HsY<S, FIRST> yA = hs_rows<S, FIRST>(state, A, iA, false, neg, mt, lane), yB = hs_rows<S, FIRST>(state, A, iB, false, neg, mt, lane);
    hs_step2<S, FIRST, TA, TB>(sfrag, lane, yA, yB, wA, wB);
    __syncwarp();
#ifndef WB_HS_PLAINBAR
    if (lane == 0) hs_mbar_arrive(barA);   // this warp has read everything it needs from other panels' rows
#endif
    if (HASB) hs_traces<S, FIRST, HASB ? TB : TA>(wB, yB, iB, needO, needE, lane, tr);
    hs_traces<S, FIRST, TA>(wA, yA, iA, needO, needE, lane, tr);
    // every panel has read its rows of B_k; the partial tail tiles are in shared memory.  (-DWB_HS_PLAINBAR: an ordinary team
    // barrier at the same place - what compute-sanitizer's racecheck can follow; it does not model mbarrier ordering.  A
    // run-time switch cost the kernel 1 %.)
#ifdef WB_HS_PLAINBAR
    S::sync(team);
#else
    hs_mbar_wait(barA, parity);
#endif
    if (store) {
        hs_store<S, TA>(state, iA, lane, wA);
        if (HASB) hs_store<S, HASB ? TB : TA>(state, iB, lane, wB);
    }
    if (TAIL && wl == 0) {
        HafRow<TF, TAIL> wC;
        wC.wtr = wC.wti = 0.0;
#pragma unroll
        for (int wv = 0; wv < S::TW; ++wv) { const double2 e = tailC[wv * 32 + lane]; wC.wtr += e.x; wC.wti += e.y; }
        hs_traces<S, FIRST, TF>(wC, yC, 4 * TF, needO, needE, lane, tr);
        __syncwarp();                      // the pairing read the tail block across lanes (t = 0 <-> 1) before it is overwritten
        if (store) hs_store<S, TF>(state, 4 * TF, lane, wC);
    }
}

template <class S, int FIRST>
__device__ __forceinline__ void hs_role_step(int rho, const double2* __restrict__ sfrag, double2* __restrict__ state, const double* __restrict__ A,
                                             int sub, bool needO, bool needE, bool store, unsigned neg, int mt,
                                             int lane, int team, int wl, double2* __restrict__ tailC, uint64_t* barA, unsigned parity, double (&tr)[6]) {
#define WB_HS_STEP(r) hs_warp_step<S, FIRST, (r) < S::ROLES ? (r) : 0>(sfrag, state, A, sub, needO, needE, store, neg, mt, lane, team, wl, tailC, barA, parity, tr)
    if (rho == 0) WB_HS_STEP(0);
    else if (rho == 1) WB_HS_STEP(1);
    else if (rho == 2 || S::ROLES == 3) WB_HS_STEP(2);
    else WB_HS_STEP(3);
#undef WB_HS_STEP
}

template <class S, bool PAD>
__global__ void __launch_bounds__(32 * S::WARPS, 1)
haf_sym_kernel(const double* __restrict__ A, uint64_t j0, uint64_t j1, double* __restrict__ partials, int m_true, long long skew_cycles) {
    // m: the TRUE number of vertex pairs (PAD: smaller than S::M, the matrix is treated as padded with zero rows / columns;
    // a run-time m costs the exact shapes 2 - 3 %, hence the separate instances)
    constexpr int TF = S::TF, PM = S::M, TW = S::TW, NQ = S::NQ;
    const int m = PAD ? m_true : PM;
    const int n = 2 * m;
    constexpr bool TAIL = S::TAIL;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int team = S::team(warp), wl = S::wl(warp);
    double2* sfrag = reinterpret_cast<double2*>(smem);
    double2* state = reinterpret_cast<double2*>(smem + S::FRAG_D) + team * NQ * S::QS;                    // [subset][row][col]
    double2* P2 = reinterpret_cast<double2*>(smem + S::FRAG_D + 2 * (size_t)S::STATE_D2) + team * NQ * (PM + 2);   // P[subset][j]
    double2* ptmp = reinterpret_cast<double2*>(smem + S::FRAG_D + 2 * (size_t)S::STATE_D2 + S::P_D) + team * 3 * TW * 8;   // [kind][warp][row]
    double2* tailC = reinterpret_cast<double2*>(smem + S::FRAG_D + 2 * (size_t)S::STATE_D2 + S::P_D + S::TMP_D) + team * TW * 32;   // [warp][lane]
    __shared__ uint64_t bars[2];
    uint64_t* barA = bars + team;
    if (threadIdx.x < S::TEAMS) hs_mbar_init(bars + threadIdx.x, TW);
    haf_build_frag(A, n, m, TF, TAIL ? 1 : 0, sfrag, threadIdx.x, 32 * S::WARPS);
    __syncthreads();
    // optional start skew of the second team (see HsShape): n = 50 gains 1 % from ~ a product's length, n = 48 loses 5 %
    if (S::TEAMS == 2 && team == 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < skew_cycles) __nanosleep(200);
    }

    const int g = lane >> 2, t = lane & 3, q = S::q(lane);
    const int rho = wl / S::SUBS, sub = wl % S::SUBS;
    const int nprod = (m - 1) >> 1, K = nprod + 1;
    const uint64_t ngroups = (j1 - j0 + NQ - 1) / NQ;

    const double inv_idx = 1.0 / (double)(lane + 1);
    unsigned uses = 0;
    cdd acc;
    acc.re = {0.0, 0.0};
    acc.im = {0.0, 0.0};

    for (uint64_t G = (uint64_t)blockIdx.x * S::TEAMS + team; G < ngroups; G += (uint64_t)gridDim.x * S::TEAMS) {
        const uint64_t jq = j0 + NQ * G + q;
        const unsigned neg = hs_negmask(jq, m);
        if (wl == 0) {
            // tr(M^1) = sum_r delta_r A'[r][sigma(r)] = 2 sum_i delta_i A'[i][i + m]; every P[1..m] is rewritten for every group
            if (lane < NQ) {
                const uint64_t jj = j0 + NQ * G + lane;
                double sr = 0.0, si = 0.0;
                for (int i = 0; i < m; ++i) {
                    const double d = ((jj >> (m - 1 - i)) & 1ull) ? 2.0 : -2.0;
                    sr += d * __ldg(A + 2 * ((size_t)i * n + i + m));
                    si += d * __ldg(A + 2 * ((size_t)i * n + i + m) + 1);
                }
                P2[lane * (PM + 2) + 1] = make_double2(sr, si);
            }
            __syncwarp();
        }
        for (int k = 1; k <= nprod; ++k) {
            const bool needO = (2 * k + 1 > K) && (2 * k + 1 <= m);
            const bool needE = (2 * k + 2 > K) && (2 * k + 2 <= m);
            double tr[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            const bool store = k < nprod;
            const unsigned parity = uses++ & 1u;            // phase of the split barrier: one use per product
            if (k == 1) hs_role_step<S, PAD ? 2 : 1>(rho, sfrag, state, A, sub, needO, needE, store, neg, m, lane, team, wl, tailC, barA, parity, tr);   // B_1 = A' from global memory
            else hs_role_step<S, 0>(rho, sfrag, state, A, sub, needO, needE, store, neg, m, lane, team, wl, tailC, barA, parity, tr);
            // per-row trace shares of this warp: reduce over the four lanes of a row, park them for the team
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                tr[c] += shfl_xor_d(tr[c], 1);
                tr[c] += shfl_xor_d(tr[c], 2);
            }
            if (t == 0) {
#pragma unroll
                for (int kind = 0; kind < 3; ++kind) ptmp[(kind * TW + wl) * 8 + g] = make_double2(tr[2 * kind], tr[2 * kind + 1]);
            }
            S::sync(team);                     // B_(k+1) is complete in shared memory; so are this step's trace shares
            if (wl < 3 * NQ) {   // warp w of the team sums the shares (TW warps x the 8 / NQ rows of a subset) of ONE (kind, subset) in a
                                 // fixed shuffle tree
                constexpr int RPS = 8 / NQ;    // rows per subset in a panel
                const int kind = wl / NQ, qq = wl % NQ;
                double2 e = make_double2(0.0, 0.0);
                if (lane < TW * RPS) e = ptmp[(kind * TW + lane / RPS) * 8 + qq + NQ * (lane % RPS)];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) { e.x += shfl_xor_d(e.x, off); e.y += shfl_xor_d(e.y, off); }
                const int j = kind == 0 ? k + 1 : (kind == 1 ? 2 * k + 1 : 2 * k + 2);
                const bool want = kind == 0 || (kind == 1 ? needO : needE);
                if (lane == 0 && want && j <= m) P2[qq * (PM + 2) + j] = e;
            }
        }
        S::sync(team);                         // the last step's traces are in P
        // ---- series c_s = (1/s) sum_i (p_i / 2) c_(s-i), first warp of the team, "push" form: lane l holds the partial sum of
        // coefficient l + 1 for each of the NQ subsets (independent chains); step s: lane s - 1 finishes c_s, one broadcast, and
        // every lane of a later coefficient c adds F_(c-s) c_s.  (The first version split each inner sum over 8 lanes with two FP64 divisions and three
        // shuffle levels per step: 28 k cycles per group, 8 % of the kernel.)
        if (wl == 0) {
            const int idx = lane + 1;                                       // this lane's coefficient index (m <= 32 fits a warp)
            double ar[NQ], ai[NQ];
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq) {
                const double2 f = idx <= m ? P2[qq * (PM + 2) + idx] : make_double2(0.0, 0.0);
                ar[qq] = 0.5 * f.x; ai[qq] = 0.5 * f.y;                    // c_0 = 1
            }
            for (int sidx = 1; sidx < m; ++sidx) {
                const bool tgt = idx > sidx && idx <= m;
#pragma unroll
                for (int qq = 0; qq < NQ; ++qq) {
                    const double cr = shfl_d(ar[qq] * inv_idx, sidx - 1), ci = shfl_d(ai[qq] * inv_idx, sidx - 1);
                    const double2 f = tgt ? P2[qq * (PM + 2) + idx - sidx] : make_double2(0.0, 0.0);
                    const double fr = 0.5 * f.x, fi = 0.5 * f.y;
                    ar[qq] = fma(fr, cr, ar[qq]); ar[qq] = fma(-fi, ci, ar[qq]);
                    ai[qq] = fma(fr, ci, ai[qq]); ai[qq] = fma(fi, cr, ai[qq]);
                }
            }
            if (idx == m) {
#pragma unroll
                for (int qq = 0; qq < NQ; ++qq) {
                    const uint64_t jj = j0 + NQ * G + qq;
                    if (jj < j1) {
                        const double sg = ((m - __popcll(jj)) & 1) ? -1.0 : 1.0;
                        dd_add(acc.re, sg * ar[qq] * inv_idx);
                        dd_add(acc.im, sg * ai[qq] * inv_idx);
                    }
                }
            }
            __syncwarp();
        }
    }
    __shared__ double red[S::WARPS * 4];
    block_reduce_store(acc, red, partials);
}

template <class S>
static int launch_haf_sym(const double* dA, int n, uint64_t j0, uint64_t j1, double* partials, int sms, int* grid_out, cudaStream_t st) {
    auto kern = (n == S::N || S::TAIL) ? haf_sym_kernel<S, false> : haf_sym_kernel<S, !S::TAIL>;   // padding: whole-tile shapes only
    const uint64_t ngroups = (j1 - j0 + 3) >> 2;
    const int grid = (int)(ngroups < (uint64_t)sms ? (ngroups ? ngroups : 1) : (uint64_t)sms);
    WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
    const char* ek = getenv("WB200_HS_SKEW");
    const long long skew = ek ? atoll(ek) : (S::TAIL ? 20000 : 0);
    kern<<<grid, 32 * S::WARPS, S::BYTES, st>>>(dA, j0, j1, partials, n / 2, skew);
    WB_CUDA(cudaGetLastError());
    *grid_out = grid;
    return WB200_OK;
}

// Used by wb200_hafnian_dev for even n in [36, 64] without loops (env WB200_HAF_SYM=0 keeps the row-panel kernel, =4 the
// one-team shape for n = 48 / 50).  Shapes: m = n / 2 = 0 or 1 (mod 4) - whole tiles of four vertex pairs plus at most one
// tail pair - that fit in shared memory: two teams of two subsets up to n = 50, one team of two subsets for n = 56 / 58 / 64;
// m = 2 or 3 (mod 4) run zero-padded in the next whole-tile shape (n = 46 as 48, 54 as 56: 1.40x / 1.43x the row-panel
// kernel; two padded pairs - 36, 44, 52 - still 1.03x / 1.13x / 1.20x).  Returns WB200_ENOSUP for other sizes.
bool haf_sym_supports(int n) { return n >= 36 && n <= 64 && (n & 1) == 0; }

int haf_sym_launch(const double* dA, int n, uint64_t j0, uint64_t j1, double* partials, int sms, int* grid_out, cudaStream_t st) {
    const char* es = getenv("WB200_HAF_SYM");
    const bool one_team = es && atoi(es) == 4;
    switch (n) {
        case 36: case 38:                      // one or two vertex pairs short of whole tiles: run padded in the next shape
        case 40: return launch_haf_sym<HsShape<5, false, 2, 2>>(dA, n, j0, j1, partials, sms, grid_out, st);
        case 42: return launch_haf_sym<HsShape<5, true, 2, 2>>(dA, n, j0, j1, partials, sms, grid_out, st);
        case 44: case 46:
        case 48: return one_team ? launch_haf_sym<HsShape<6, false, 4, 1>>(dA, n, j0, j1, partials, sms, grid_out, st)
                                 : launch_haf_sym<HsShape<6, false, 2, 2>>(dA, n, j0, j1, partials, sms, grid_out, st);
        case 50: return one_team ? launch_haf_sym<HsShape<6, true, 4, 1>>(dA, n, j0, j1, partials, sms, grid_out, st)
                                 : launch_haf_sym<HsShape<6, true, 2, 2>>(dA, n, j0, j1, partials, sms, grid_out, st);
        case 52: case 54:
        case 56: return launch_haf_sym<HsShape<7, false, 2, 1>>(dA, n, j0, j1, partials, sms, grid_out, st);
        case 58: return launch_haf_sym<HsShape<7, true, 2, 1>>(dA, n, j0, j1, partials, sms, grid_out, st);
        case 60: case 62:
        case 64: return launch_haf_sym<HsShape<8, false, 2, 1>>(dA, n, j0, j1, partials, sms, grid_out, st);
        default: return WB200_ENOSUP;
    }
}

}  // namespace wb
