// Hafnian / loop hafnian, all edge repetitions 1, Glynn sieve — FP64 tensor-core (DMMA.8x8x4) kernel.
//
// Replaces the prange body of _calc_hafnian / _calc_loop_hafnian (thewalrus/_hafnian.py:416-467,
// 512-577) + charpoly.powertrace (thewalrus/charpoly.py:301-327) + f / f_loop (_hafnian.py:183-242)
// for edge_reps = [1]*m, glynn=True.
//
// Math: with A' in matched order (vertex e paired with sigma(e) = e +- m), subset
// j <-> delta in {+-1}^m, S_j = X diag(delta, delta), M_j = A' S_j.  B_k := M_j^k S_j is symmetric,
// B_1 = A', and row c of B_{k+1} = (row c of B_k) S_j A'.  Rows evolve independently, so a warp keeps
// 8 rows (4 subsets x the vertex pair {i, i+m}) in registers as DMMA A-operand fragments and multiplies
// by the fixed matrix A' whose B-operand fragments are staged once per CTA in shared memory.  The
// D-fragment (accumulator) register layout of m8n8k4 is reused directly as the next A-fragment by
// choosing the K-chunk <-> column mapping accordingly, so the chain needs no shuffles or smem traffic
// for the iterate.  Power traces:
//     tr(M^k)     = sum_c delta_c B_k[c, sigma(c)]                          (one element per row: free)
//     tr(M^(a+b)) = sum_c delta_c <row c of B_a, S_j (row sigma(c) of B_b)>  (pairing, for k > K)
// so only K - 1 = ceil(m/2) - 1 products are needed for tr(M^1..M^m) (the reference does m-1,
// charpoly.py:316-318).  Loop hafnian: one more register row per subset, Z_k = (M^k D)^T, with
//     XD^T M^(a+b) D = <Z_a, S_j Z_b>            (reference: _hafnian.py:233-234).
//
// Register layout of a row for thread (g = lane >> 2, t = lane & 3):
//   full tile tau < TF : w[tau][r] = element 4 tau + t + r m   (partners share a thread: swap is free)
//   tail tile (TAIL)   : (wt_r, wt_i) = element 4 TF + (t >> 1) + (t & 1) m, partner in lane ^ 1.
// The tail tile packs the last m mod 4 in {1, 2} vertex pairs with (re, im) interleaved so that sizes
// like n = 50 (m = 25) waste one quarter of one tile instead of a whole tile row and column.  With ONE pair in
// the tail (m mod 4 = 1) the tail K-chunk is packed as well — positions 2, 3 carry the imaginary parts — and
// costs two DMMAs per tile instead of four (haf_prep_kernel / haf_step, `packk`).
// Launch shape: one CTA per SM, 12 warps for full-size problems, 4 or 8 warps when there are fewer groups of
// four subsets than that, so small problems spread over all SMs (haf_pick_warps).
#include <stdlib.h>
#include "common.cuh"
#include "haf_dmma.cuh"

namespace wb {

template <int TF, bool TAIL, int WARPS>
struct HafCfg {
    static constexpr int NK = 2 * TF + (TAIL ? 1 : 0);
    static constexpr int NT = TF + (TAIL ? 1 : 0);
    static constexpr int MP = 4 * TF + (TAIL ? 2 : 0);     // upper bound on m
    static constexpr int FRAG_D = NK * NT * 64;            // doubles in fragment table
    static constexpr int PART_D = (MP + 2) * 8 * 2;        // part[j][g] complex, j = 0..MP+1
    static constexpr int P_D = (MP + 2) * 4 * 2;           // P[k][q] complex
    static constexpr int WARP_D = PART_D + 3 * P_D;        // partials, P, L, c
    static constexpr size_t BYTES = sizeof(double) * (FRAG_D + WARPS * WARP_D);
};

// PS = panel split: the row panels of one group of four subsets are dealt round-robin to a TEAM of PS warps (on
// different schedulers), whose per-panel trace shares are combined in fixed warp order by the team's first warp.
// PS = 1 (one warp walks all panels) is the full-size shape; PS = 3 serves small problems (n <= 30), where there are
// too few groups to give every scheduler three independent warps (n = 24: 512 groups for 592 schedulers).
template <int TF, bool TAIL, int WARPS, int PS>
__global__ void __launch_bounds__(32 * WARPS, 1)
haf_dmma_kernel(const double* __restrict__ A, const double* __restrict__ D, int n,
                int m, uint64_t j0, uint64_t j1, double* __restrict__ partials) {
    using C = HafCfg<TF, TAIL, WARPS>;
    static_assert(WARPS % PS == 0, "teams must tile the CTA");
    extern __shared__ __align__(16) double smem[];
    double2* sfrag = reinterpret_cast<double2*>(smem);
    haf_build_frag(A, n, m, TF, TAIL ? 1 : 0, sfrag, threadIdx.x, 32 * WARPS);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3, q = g & 3, half = g >> 2;
    double* wsm = smem + C::FRAG_D + warp * C::WARP_D;
    double2* part = reinterpret_cast<double2*>(wsm);  // part[j * 8 + g]: row-g share of tr(M^j)
    double* Pk = wsm + C::PART_D;                     // P[k][q] complex: Pk[(k*4+q)*2 + {0,1}]
    double* Lk = Pk + C::P_D;                         // loop terms
    double* Ck = Lk + C::P_D;                         // series coefficients
    const bool loop = (D != nullptr);
    const int tp = TAIL ? m - 4 * TF : 0;             // vertex pairs in the tail tile (1 or 2)
    const int nprod = (m - 1) >> 1;                   // products: B_2 .. B_K, K = nprod + 1
    const int K = nprod + 1;                          // tr(M^j), j <= K, come from single elements
    const int nstepD = m >> 1;                        // products of the loop row (l_1..l_m)
    const uint64_t ngroups = (j1 - j0 + 3) >> 2;
    constexpr int TEAMS = WARPS / PS;
    const int team = warp / PS, pw = warp - team * PS;       // warp pw of its team takes panels pw, pw + PS, ...
    const uint64_t gstride = (uint64_t)gridDim.x * TEAMS;
    auto team_barrier = [&]() {
        if constexpr (PS > 1) asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(32 * PS) : "memory");
    };

    cdd acc;
    acc.re = {0.0, 0.0};
    acc.im = {0.0, 0.0};

    for (uint64_t G = (uint64_t)blockIdx.x * TEAMS + team; G < ngroups; G += gstride) {
        const uint64_t jq = j0 + 4 * G + q;
        const bool valid = jq < j1;
        unsigned sm[TF > 0 ? TF : 1];
#pragma unroll
        for (int tau = 0; tau < TF; ++tau) {
            const int i = 4 * tau + t;
            const unsigned kept = (i < m) ? (unsigned)((jq >> (m - 1 - i)) & 1ull) : 1u;
            sm[tau] = kept ? 0u : 0x80000000u;
        }
        unsigned smt = 0u;
        if (TAIL && (t >> 1) < tp) smt = ((jq >> (m - 1 - (4 * TF + (t >> 1)))) & 1ull) ? 0u : 0x80000000u;
        for (int s = lane; s < (m + 2) * 8; s += 32) part[s] = make_double2(0.0, 0.0);
        __syncwarp();

        const int npanels = m + (loop ? 1 : 0);
        for (int i = pw; i < npanels; i += PS) {
            const bool isD = (i == m);
            HafRow<TF, TAIL> w;
            HafY<TF, TAIL> y = {};
            // ---- first iterate: row v of A' (B_1 = A'), or D for the loop row
            {
                const int v = i + half * m;
                const double* src = isD ? D : (A + 2 * (size_t)v * n);
                const bool rowok = isD ? (half == 0) : true;
#pragma unroll
                for (int tau = 0; tau < TF; ++tau) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int iv = 4 * tau + t;
                        const bool ok = rowok && (iv < m);
                        const int e = iv + r * m;
                        w.wr[tau][r] = ok ? __ldg(src + 2 * e) : 0.0;
                        w.wi[tau][r] = ok ? __ldg(src + 2 * e + 1) : 0.0;
                    }
                }
                w.wtr = w.wti = 0.0;
                if (TAIL) {
                    const bool ok = rowok && ((t >> 1) < tp);
                    const int e = 4 * TF + (t >> 1) + (t & 1) * m;
                    w.wtr = ok ? __ldg(src + 2 * e) : 0.0;
                    w.wti = ok ? __ldg(src + 2 * e + 1) : 0.0;
                }
            }
            // delta of this row's vertex pair in subset jq
            const double rs = isD ? 1.0 : (((jq >> (m - 1 - (isD ? 0 : i))) & 1ull) ? 1.0 : -1.0);
            // which thread of the row holds element sigma(v) (the "diagonal" of B_k S), and where
            const bool in_tail = TAIL && (i >= 4 * TF);
            const int own_t = in_tail ? 2 * (i - 4 * TF) + (1 - half) : (i & 3);
            const int own_tau = i >> 2;
            double orr, oi, er, ei;
            if (!isD) {
                // tr(M) share: A'[v, sigma(v)]; tr(M^2) only by pairing when K < 2 (m <= 2)
                if (K < 2) haf_advance<TF, TAIL, true, false>(w, y, sm, smt, orr, oi, er, ei);
                else haf_advance<TF, TAIL, false, false>(w, y, sm, smt, orr, oi, er, ei);
                if (t == 0) {
                    const int v = i + half * m, sv = i + (1 - half) * m;
                    double2 p = part[1 * 8 + g];
                    p.x += rs * __ldg(A + 2 * ((size_t)v * n + sv));
                    p.y += rs * __ldg(A + 2 * ((size_t)v * n + sv) + 1);
                    part[1 * 8 + g] = p;
                    if (K < 2 && m >= 2) {
                        double2 p2 = part[2 * 8 + g];
                        p2.x += rs * er; p2.y += rs * ei;
                        part[2 * 8 + g] = p2;
                    }
                }
            } else {
                HafY<TF, TAIL> y0;  // l_1 = <Z_0, S Z_0>
                haf_advance<TF, TAIL, false, true>(w, y0, sm, smt, orr, oi, er, ei);
                y = y0;
                haf_advance<TF, TAIL, true, true>(w, y, sm, smt, orr, oi, er, ei);
                if (t == 0 && half == 0) { Lk[(1 * 4 + q) * 2] = orr; Lk[(1 * 4 + q) * 2 + 1] = oi; }
            }
            const int nsteps = isD ? nstepD : nprod;
            for (int k = 1; k <= nsteps; ++k) {
                haf_step<TF, TAIL>(sfrag, lane, y, w, TAIL && tp == 1);  // w = row of B_{k+1} (Z_k for the loop row)
                if (!isD) {
                    // tr(M^(k+1)) share: element sigma(v) of this row
                    if (t == own_t) {
                        double dr, di;
                        if (in_tail) { dr = w.wtr; di = w.wti; }
                        else {
                            dr = 0.0; di = 0.0;
#pragma unroll
                            for (int tau = 0; tau < TF; ++tau)
                                if (tau == own_tau) { dr = half ? w.wr[tau][0] : w.wr[tau][1]; di = half ? w.wi[tau][0] : w.wi[tau][1]; }
                        }
                        double2 p = part[(k + 1) * 8 + g];
                        p.x += rs * dr; p.y += rs * di;
                        part[(k + 1) * 8 + g] = p;
                    }
                    const bool needO = (2 * k + 1 > K) && (2 * k + 1 <= m);
                    const bool needE = (2 * k + 2 > K) && (2 * k + 2 <= m);
                    if (needO || needE) {
                        haf_advance<TF, TAIL, true, false>(w, y, sm, smt, orr, oi, er, ei);
                        if (t == 0) {
                            if (needO) { double2 p = part[(2 * k + 1) * 8 + g]; p.x += rs * orr; p.y += rs * oi; part[(2 * k + 1) * 8 + g] = p; }
                            if (needE) { double2 p = part[(2 * k + 2) * 8 + g]; p.x += rs * er; p.y += rs * ei; part[(2 * k + 2) * 8 + g] = p; }
                        }
                    } else {
                        haf_advance<TF, TAIL, false, false>(w, y, sm, smt, orr, oi, er, ei);
                    }
                } else {  // l_{2k} = <Z_k, S Z_{k-1}>, l_{2k+1} = <Z_k, S Z_k>
                    haf_advance<TF, TAIL, true, true>(w, y, sm, smt, orr, oi, er, ei);
                    if (t == 0 && half == 0) {
                        Lk[((2 * k) * 4 + q) * 2] = orr; Lk[((2 * k) * 4 + q) * 2 + 1] = oi;
                        Lk[((2 * k + 1) * 4 + q) * 2] = er; Lk[((2 * k + 1) * 4 + q) * 2 + 1] = ei;
                    }
                }
            }
            __syncwarp();  // part[] slots are updated by different lanes in the next panel
        }
        // ---- combine the two rows (vertex i and i + m) of each subset q (and, for PS > 1, the team's warps in order)
        team_barrier();                                             // (A) every panel of the group is done
        if (PS > 1 && pw != 0) { team_barrier(); continue; }       // (B) below; the team's first warp finishes the group
        for (int s = lane; s < (m + 1) * 4; s += 32) {
            const int j = s >> 2, qq = s & 3;
            double pr = 0.0, pi = 0.0;
#pragma unroll
            for (int tw = 0; tw < PS; ++tw) {
                const double2* pt = reinterpret_cast<const double2*>(wsm + tw * C::WARP_D);
                const double2 a = pt[j * 8 + qq], b = pt[j * 8 + qq + 4];
                pr += a.x + b.x; pi += a.y + b.y;
            }
            Pk[(j * 4 + qq) * 2] = pr; Pk[(j * 4 + qq) * 2 + 1] = pi;
        }
        if (PS > 1 && loop && (m % PS) != 0) {                      // loop terms live with the warp that walked the loop row
            const double* Lsrc = wsm + (m % PS) * C::WARP_D + C::PART_D + C::P_D;
            for (int s = lane; s < (m + 2) * 8; s += 32) Lk[s] = Lsrc[s];
        }
        team_barrier();                                             // (B) the other warps may reuse their slices
        __syncwarp();
        // ---- coefficient [eta^m] of exp(sum_i a_i eta^i), a_i = p_i/(2i) (+ l_i/2): c_t = (1/t) sum_i i a_i c_{t-i}
        if (t == 0 && half == 0) {
            Ck[(0 * 4 + q) * 2] = 1.0; Ck[(0 * 4 + q) * 2 + 1] = 0.0;
            for (int tt = 1; tt <= m; ++tt) {
                double sr = 0.0, si = 0.0;
                for (int i = 1; i <= tt; ++i) {
                    double fr = 0.5 * Pk[(i * 4 + q) * 2], fi = 0.5 * Pk[(i * 4 + q) * 2 + 1];
                    if (loop) { fr += 0.5 * i * Lk[(i * 4 + q) * 2]; fi += 0.5 * i * Lk[(i * 4 + q) * 2 + 1]; }
                    const double c_r = Ck[((tt - i) * 4 + q) * 2], c_i = Ck[((tt - i) * 4 + q) * 2 + 1];
                    sr = fma(fr, c_r, sr); sr = fma(-fi, c_i, sr);
                    si = fma(fr, c_i, si); si = fma(fi, c_r, si);
                }
                Ck[(tt * 4 + q) * 2] = sr / tt; Ck[(tt * 4 + q) * 2 + 1] = si / tt;
            }
            if (valid) {
                // prefactor (-1)^(m - sum kept) (_hafnian.py:456); kept_0 = 0 always under Glynn halving
                const int nk = __popcll(jq);
                const double sg = ((m - nk) & 1) ? -1.0 : 1.0;
                dd_add(acc.re, sg * Ck[(m * 4 + q) * 2]);
                dd_add(acc.im, sg * Ck[(m * 4 + q) * 2 + 1]);
            }
        }
        __syncwarp();
    }
    __shared__ double red[WARPS * 4];
    block_reduce_store(acc, red, partials);
}

// Launch shape (one CTA per SM).  Full-size problems: 12 warps (<= 168 registers, 3 warps per scheduler), each warp
// walks all row panels of its group — measured on B200: 12 warps beat 8 by 2-4 % at n = 40..64
// (profiles/r01_hafnian_sweep.txt).  Small problems (n <= 30, i.e. at most 4096 groups of four subsets) do not have
// three independent groups per scheduler: there the panels of a group are split over a team of 3 warps
// (PS = 3; 4 teams per CTA, or one team per CTA when there are fewer groups than SMs), so that every scheduler
// still interleaves 3 warps and the launch takes ceil(groups / (4 * SMs)) short rounds instead of one or two long ones.
// env WB200_HAF_WARPS = 4 | 8 | 12 forces the round-1 shapes (PS = 1) for A/B measurements.
static int haf_env_warps() {
    const char* e = getenv("WB200_HAF_WARPS");
    const int v = e ? atoi(e) : 0;
    return (v == 4 || v == 8 || v == 12) ? v : 0;
}

template <int TF, bool TAIL, int WARPS, int PS>
static int launch_haf_w(const double* dA, const double* dD, int n, int m, uint64_t j0, uint64_t j1,
                        double* partials, uint64_t ngroups, int sms, int* grid_out, cudaStream_t st) {
    using C = HafCfg<TF, TAIL, WARPS>;
    auto kern = haf_dmma_kernel<TF, TAIL, WARPS, PS>;
    constexpr int TEAMS = WARPS / PS;
    uint64_t want = (ngroups + TEAMS - 1) / TEAMS;
    const int grid = (int)(want < (uint64_t)sms ? (want ? want : 1) : (uint64_t)sms);
    WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
    kern<<<grid, 32 * WARPS, C::BYTES, st>>>(dA, dD, n, m, j0, j1, partials);
    WB_CUDA(cudaGetLastError());
    *grid_out = grid;
    return WB200_OK;
}

template <int TF, bool TAIL>
static int launch_haf(const double* dA, const double* dD, int n, int m, uint64_t j0, uint64_t j1,
                      double* partials, uint64_t ngroups, int sms, int* grid_out, cudaStream_t st) {
    const int env = haf_env_warps();
    if constexpr (TF <= 4) {      // n <= 32: the panel-split shapes exist
        if (!env && m <= 15) {
            if (ngroups <= (uint64_t)sms) return launch_haf_w<TF, TAIL, 3, 3>(dA, dD, n, m, j0, j1, partials, ngroups, sms, grid_out, st);
            return launch_haf_w<TF, TAIL, 12, 3>(dA, dD, n, m, j0, j1, partials, ngroups, sms, grid_out, st);
        }
    }
    int warps = env;
    if (!warps) warps = ngroups <= 4ull * (uint64_t)sms ? 4 : (ngroups <= 8ull * (uint64_t)sms ? 8 : 12);
    if (warps == 12) return launch_haf_w<TF, TAIL, 12, 1>(dA, dD, n, m, j0, j1, partials, ngroups, sms, grid_out, st);
    if (warps == 8) return launch_haf_w<TF, TAIL, 8, 1>(dA, dD, n, m, j0, j1, partials, ngroups, sms, grid_out, st);
    return launch_haf_w<TF, TAIL, 4, 1>(dA, dD, n, m, j0, j1, partials, ngroups, sms, grid_out, st);
}

constexpr int HAF_MAX_GRID = 4096;

// hafnian_sym.cu: symmetric-half kernel for n = 48 / 50 / 56 (WB200_ENOSUP for other shapes)
bool haf_sym_supports(int n);
int haf_sym_launch(const double* dA, int n, uint64_t j0, uint64_t j1, double* partials, int sms, int* grid_out, cudaStream_t st);

}  // namespace wb

using namespace wb;

extern "C" size_t wb200_hafnian_workspace_bytes(int n) {
    if (n < 2 || n > 64 || (n & 1)) return 0;
    return sizeof(double) * ((size_t)17 * 9 * 64 + (size_t)HAF_MAX_GRID * 4) + 256;
}

extern "C" int wb200_hafnian_dev(const double* dA, const double* dD, int n, uint64_t j0, uint64_t j1, double* d_out4,
                                 void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!dA || !d_out4 || !d_workspace) { set_error("hafnian: null pointer"); return WB200_EINVAL; }
    if (n < 2 || (n & 1)) { set_error("hafnian: n must be even and >= 2 (got %d)", n); return WB200_EINVAL; }
    if (n > 64) { set_error("hafnian: n = %d exceeds the DMMA kernel limit of 64", n); return WB200_ENOSUP; }
    const int m = n / 2;
    int TF, tail;
    haf_shape(m, &TF, &tail);
    const uint64_t steps = 1ull << (m - 1);
    if (j0 > j1 || j1 > steps) { set_error("hafnian: bad subset range"); return WB200_EINVAL; }
    if (workspace_bytes < wb200_hafnian_workspace_bytes(n)) { set_error("hafnian: workspace too small"); return WB200_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    (void)cudaGetLastError();  // drop any stale non-sticky error from earlier calls
    WB_CUDA(cudaGetDevice(&dev));
    if (device_sm_count(dev, &sms)) return WB200_ECUDA;
    double* partials = reinterpret_cast<double*>(d_workspace) + (size_t)17 * 9 * 64;   // (the table slot of round 1 is unused)
    const uint64_t ngroups = (j1 - j0 + 3) >> 2;
    int grid = 1;
    int rc = WB200_ENOSUP;
    {   // full-size hafnians of n = 48 / 50 / 56 without loops: only the tiles on and above the diagonal of every product
        const char* es = getenv("WB200_HAF_SYM");
        const bool want = !(es && atoi(es) == 0);
        if (want && !dD && haf_sym_supports(n) && ngroups >= 8ull * (uint64_t)sms) {
            rc = haf_sym_launch(dA, n, j0, j1, partials, sms, &grid, st);
            if (rc) return rc;
            final_reduce_kernel<<<1, 32, 0, st>>>(partials, grid, d_out4);
            WB_CUDA(cudaGetLastError());
            return WB200_OK;
        }
    }
#define WB_HAF_CASE(tf, tl) case (tf) * 2 + (tl): rc = launch_haf<tf, (tl) != 0>(dA, dD, n, m, j0, j1, partials, ngroups, sms, &grid, st); break;
    switch (TF * 2 + tail) {
        WB_HAF_CASE(1, 0) WB_HAF_CASE(2, 0) WB_HAF_CASE(3, 0) WB_HAF_CASE(4, 0)
        WB_HAF_CASE(5, 0) WB_HAF_CASE(6, 0) WB_HAF_CASE(7, 0) WB_HAF_CASE(8, 0)
        WB_HAF_CASE(1, 1) WB_HAF_CASE(2, 1) WB_HAF_CASE(3, 1) WB_HAF_CASE(4, 1)
        WB_HAF_CASE(5, 1) WB_HAF_CASE(6, 1) WB_HAF_CASE(7, 1)
        default: set_error("hafnian: unsupported tile shape TF=%d tail=%d", TF, tail);
    }
#undef WB_HAF_CASE
    if (rc) return rc;
    final_reduce_kernel<<<1, 32, 0, st>>>(partials, grid, d_out4);
    WB_CUDA(cudaGetLastError());
    return WB200_OK;
}
