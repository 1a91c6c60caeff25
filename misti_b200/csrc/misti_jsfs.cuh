// misti_jsfs.cuh -- expected joint SFS of one evaluation item, written for a cooperating GROUP of
// lanes (a warp on the device; a single "lane" in the test-only host build tests/hostsim).
//
// Reference path restated here: MigrationInference.JAFSpectrum / SolveDifEq / CollapsePops
// (MigrationInference.py:467-540), TwoPopulations.SetMatrix / UpdateMatrixCol / PulseMigration /
// AncientSampleP0 / StateToJAF (TwoPopulations.py:188-262, 336-377), OnePopulation.SetMatrix
// (OnePopulation.py:153-178), and the likelihood tail (MigrationInference.py:583-613).
//
// What is different from the reference (same numbers, different algorithm):
//   * P1 = expm(M T) P0 and integralP = inv(M)(P1 - P0) = int_0^T exp(M s) P0 ds are obtained
//     together by UNIFORMISATION of the lineage chain: with q >= max |M_cc| the matrix
//     A = I + M/q is non-negative, exp(M T) = sum_k Pois(k; qT) A^k, and
//     int_0^T exp(M s) P0 ds = (1/q) sum_k Pois(k; qT) (P0 + A P0 + ... + A^(k-1) P0).
//     Every term is non-negative (no cancellation), only sparse 44x44 mat-vecs are needed
//     (152 off-diagonal entries), the zero-migration singular case of the reference
//     (TwoPopulations.py:240-309, 7 stationary states removed and patched back) needs no special
//     handling, and no inverse is formed.  Intervals with qT > 32 are cut into equal sub-steps.
//   * after the split all generators are multiples of one constant 8x8 matrix L8 and commute, so
//     the whole post-split contribution is  sum_k cpost[k] * (W8 G_k) P8  with the three spectral
//     projectors G_k of L8 (eigenvalues -6, -3, -1); cpost[] is accumulated by post_split_coeffs().
#pragma once
#include "misti_model.cuh"
#include "misti_tables.h"

namespace misti {

struct EllEntry { unsigned char col, kind, cnt; };
struct PulseEntry { unsigned char row, col, a, b, mult; };

#define MISTI_DEFINE_TABLES(SPEC, PFX)                                          \
    SPEC EllEntry PFX##ell[44][MISTI_ELL_WIDTH] = MISTI_ELL_INIT;               \
    SPEC unsigned char PFX##diag[44][4] = MISTI_GEN_DIAG_INIT;                  \
    SPEC unsigned char PFX##w44[7][44] = MISTI_W44_INIT;                        \
    SPEC unsigned char PFX##collapse[44] = MISTI_COLLAPSE_INIT;                 \
    SPEC unsigned char PFX##anc2[44] = MISTI_ANC2_INIT;                         \
    SPEC unsigned char PFX##anc11[44] = MISTI_ANC11_INIT;                       \
    SPEC PulseEntry PFX##pulse0[MISTI_PULSE0_NNZ] = MISTI_PULSE0_INIT;          \
    SPEC PulseEntry PFX##pulse1[MISTI_PULSE1_NNZ] = MISTI_PULSE1_INIT;          \
    SPEC unsigned char PFX##pulse0_rowptr[45] = MISTI_PULSE0_ROWPTR_INIT;       \
    SPEC unsigned char PFX##pulse1_rowptr[45] = MISTI_PULSE1_ROWPTR_INIT;       \
    SPEC double PFX##wg6[7][8] = MISTI_WG6_INIT;                                \
    SPEC double PFX##wg3[7][8] = MISTI_WG3_INIT;                                \
    SPEC double PFX##wg1[7][8] = MISTI_WG1_INIT;

// Under nvcc the table users are device-only functions reading __device__ copies; the test-only
// host build (g++) reads plain static copies.
#if defined(__CUDACC__)
MISTI_DEFINE_TABLES(static __device__ const, d_)
#define MISTI_TAB(name) d_##name
#define MISTI_D __device__
#else
MISTI_DEFINE_TABLES(static const, h_)
#define MISTI_TAB(name) h_##name
#define MISTI_D
#endif

// ---- lane groups ------------------------------------------------------------------------------
// A group = the lanes that share one item.  Several groups may share a warp; they then run in LOCK
// STEP: every loop bound and branch in jsfs_item is made warp-uniform with wmax / any / all, and a
// group that has nothing left to do keeps executing on a zero-length interval (an exact no-op).
struct SingleLane {  // test-only host build: one lane owns all 44 rows
    static constexpr int LANES = 1;
    MISTI_HD int lane() const { return 0; }
    MISTI_HD void sync() const {}
    MISTI_HD double max(double v) const { return v; }
    MISTI_HD double sum(double v) const { return v; }
    MISTI_HD int wmax(int v) const { return v; }
    MISTI_HD bool any(bool v) const { return v; }
    MISTI_HD bool all(bool v) const { return v; }
    // coefficient table: entry c < 12 is 2^(c >> 2) * rq[c & 3], entries 12.. are 0
    MISTI_HD double table_entry(const double*) const { return 0.0; }
    MISTI_HD double table_get(double, unsigned c, const double* rq) const {
        return c >= 12u ? 0.0 : (double)(1u << (c >> 2)) * rq[c & 3u];
    }
};

#if defined(__CUDACC__)
struct HalfWarpLanes {  // two items per warp: lane l of each 16-lane half owns rows l, l + 16, l + 32
    static constexpr int LANES = 16;
    __device__ int lane() const { return threadIdx.x & 15; }
    __device__ void sync() const { __syncwarp(); }
    __device__ double max(double v) const {  // within the half
        for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ double sum(double v) const {  // within the half
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ int wmax(int v) const {  // over the whole warp (both items); v is uniform within a half
        const int o = __shfl_xor_sync(0xffffffffu, v, 16);
        return v > o ? v : o;
    }
    __device__ bool any(bool v) const { return __any_sync(0xffffffffu, v); }
    __device__ bool all(bool v) const { return __all_sync(0xffffffffu, v); }
    // coefficient table spread over the 16 lanes of the group: lane c < 12 holds 2^(c >> 2) * rq[c & 3],
    // lanes 12.. hold 0; table_get fetches entry (c & 15) with one shuffle instead of a select chain
    __device__ double table_entry(const double* rq) const {
        const unsigned c = threadIdx.x & 15u;
        const unsigned kind = c & 3u;
        const double base = kind == 0 ? rq[0] : (kind == 1 ? rq[1] : (kind == 2 ? rq[2] : rq[3]));
        return c >= 12u ? 0.0 : (double)(1u << (c >> 2)) * base;
    }
    __device__ double table_get(double mine, unsigned c, const double*) const { return __shfl_sync(0xffffffffu, mine, c, 16); }
};
#endif

// 1/k for the Poisson weight recursion p_k = p_(k-1) * lam / k (a table lookup instead of a division per term)
#define MISTI_RECIP_N 512
struct RecipTable {
    double v[MISTI_RECIP_N];
    constexpr RecipTable() : v() {
        v[0] = 0.0;
        for (int k = 1; k < MISTI_RECIP_N; ++k) v[k] = 1.0 / k;
    }
};
#if defined(__CUDACC__)
static __constant__ RecipTable c_recip = RecipTable();
#define MISTI_RECIP(k) c_recip.v[k]
#else
static const RecipTable h_recip = RecipTable();
#define MISTI_RECIP(k) h_recip.v[k]
#endif

constexpr int kYStride = 48;                    // doubles per ping-pong buffer (44 states + pad rows)
constexpr double kUnifMaxStep = 32.0;           // largest q*T handled in one uniformisation sweep
constexpr double kUnifTol = 1.3877787807814457e-17;  // 2^-56: truncation of the Poisson tail
constexpr double kUnifMaxStiff = 256.0;         // q*T beyond this (8 sweeps) goes to the dense scaling-and-squaring step

// Continuation record of an item whose next two-population interval is too stiff for the uniformisation sweep
// (rates ~1e5..1e8 after a run-away correction): the state at the START of interval `it` (before the ancient
// reset / pulse of that interval).  misti_stiff_kernel advances it over the stiff interval(s) with a dense
// scaling-and-squaring step; the JSFS kernel then resumes from it.
struct Cont {
    int it;
    int nterms;
    int pad_[2];
    double P[48];
    double Ia[48];
    double Ib[48];
};
constexpr int kUnifMaxTerms = MISTI_RECIP_N - 4;     // never reached for q*T <= 32 (about 110 terms)

// Post-split coefficients (run by ONE thread; lc addressed like in correct_lambdas_item):
//   cpost[k] = sum_{i>=splitT} exp(-a_k x_i) (1 - exp(-a_k lam_i T_i)) / (a_k lam_i),  x_i = sum_{j<i} lam_j T_j,
// with the last interval infinite (MigrationInference.py:530-540: P1 = 0 there), a = (6, 3, 1).
MISTI_HD inline void post_split_coeffs(const ModelDesc& md, const double* times, const double* lc, int pitch, long stride,
                                       double* cpost) {
    double c6 = 0, c3 = 0, c1 = 0;
    double e1 = 1.0;  // exp(-x)
    for (int t = md.splitT; t < md.numT; ++t) {
        const double lam = lc[(pitch * t) * stride];
        const double e3 = e1 * e1 * e1, e6 = e3 * e3;
        if (t < md.numT - 1) {
            const double z = lam * times[t];
            const double u = exp(-z), w1 = -expm1(-z);      // w1 = 1 - u
            const double w3 = w1 * (1.0 + u + u * u);         // 1 - u^3
            const double w6 = w3 * (1.0 + u * u * u);         // 1 - u^6
            c1 += e1 * w1 / lam;
            c3 += e3 * w3 / (3.0 * lam);
            c6 += e6 * w6 / (6.0 * lam);
            e1 *= u;
        } else {
            c1 += e1 / lam;
            c3 += e3 / (3.0 * lam);
            c6 += e6 / (6.0 * lam);
        }
    }
    cpost[0] = c6; cpost[1] = c3; cpost[2] = c1;
}

// Rates are read as lc[(PITCH*t + j)*stride]: j = 0, 1 the corrected coalescence rates and, when PITCH == 4, j = 2, 3
// the migration rates of the interval (written by the correction kernel; with PITCH == 2 they are re-derived from the
// band list).
// Expected JSFS of one item.  ALL lanes of the warp call this together (each group with its own item;
// `active` = false for a group without work); `ysm` is a scratch area of 2*kYStride doubles private to the
// group (two ping-pong copies of the 44-vector, padded to 48 so that every lane has a slot to write).  On return every lane of the group holds the UNNORMALISED spectrum in jafs[0..6]
// (MigrationInference.JAFSpectrum's return value) and the number of mat-vecs in *terms.
// `cont` (nullable): where to park the item when it meets a stiff interval (return value MISTI_STIFF = "pending");
// `resume`: start from the record instead of from the sampling configuration.
template <class G, int PITCH>
MISTI_D inline int jsfs_item(const G& g, const ModelDesc& md, bool active, const double* times, const double* params,
                             const double* lc, long stride, const double* cpost, double* ysm, double* jafs, int* terms,
                             Cont* cont = nullptr, bool resume = false) {
    constexpr int RPL = (44 + G::LANES - 1) / G::LANES;
    constexpr int W = MISTI_ELL_WIDTH;
    const int lane = g.lane();
    bool valid[RPL];
    // per row, packed: ELL slot e -> coefficient-table code (kind + 4 log2(count), 12 = empty) at bit 4e;
    // diagonal multiplicity of rate kind k at bit 16+3k
    unsigned rc[RPL];
    const double* yp[RPL][W];  // where this lane reads y[col] for each ELL slot (buffer 0; buffer 1 is +kYStride)
    double* const wb = ysm + lane;  // this lane writes y[row] at wb[s * LANES] (+kYStride for buffer 1); rows 44.. are pad
    double P[RPL];            // state probabilities at the start of the current interval (rows owned by this lane)
    double Ia[RPL], Ib[RPL];  // occupancy integrals summed over the intervals before / from the sampling date
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int r = lane + s * G::LANES;
        valid[s] = r < 44;
        const int rr = valid[s] ? r : 0;
        rc[s] = 0;
#pragma unroll
        for (int e = 0; e < W; ++e) {
            const EllEntry en = MISTI_TAB(ell)[rr][e];
            const unsigned cde = (!valid[s] || en.cnt == 0) ? 12u : (unsigned)en.kind + (en.cnt == 1 ? 0u : (en.cnt == 2 ? 4u : 8u));
            rc[s] |= cde << (4 * e);
            yp[s][e] = ysm + (valid[s] ? en.col : 0);
        }
        if (valid[s])
            for (int k = 0; k < 4; ++k) rc[s] |= (unsigned)MISTI_TAB(diag)[r][k] << (16 + 3 * k);
        P[s] = (valid[s] && r == 2) ? 1.0 : 0.0;  // both genome-1 lineages in deme 0, genome-2 in deme 1 (:469-471)
        Ia[s] = 0.0; Ib[s] = 0.0;
    }
    int nterms = 0;
    int status = MISTI_OK;
    int it0 = 0;
    if (resume && active && cont) {
        it0 = cont->it;
        nterms = cont->nterms;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int r = lane + s * G::LANES;
            P[s] = valid[s] ? cont->P[r] : 0.0;
            Ia[s] = valid[s] ? cont->Ia[r] : 0.0;
            Ib[s] = valid[s] ? cont->Ib[r] : 0.0;
        }
    }
    bool pending = false;
    const int numT = md.numT;
    const int n2 = !active ? 0 : (md.splitT < numT ? md.splitT : numT);  // two-population intervals of this item
    const bool inf_last = active && md.splitT >= numT;                    // ... the last of which is then infinite
    const int n_fin = inf_last ? n2 - 1 : n2;
    const bool has_pulses = md.n_pulses > 0;

    // AncientSampleP0 (TwoPopulations.py:246-262) at it == sampleDate, then PulseMigration (:361-377)
    auto reset_and_pulse = [&](int it, bool act) {
        const bool do_reset = act && it == md.sampleDate && it > 0;  // identity on the start vector
        if (g.any(do_reset)) {
            double a2 = 0.0, a11 = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s]) {
                    if (MISTI_TAB(anc2)[lane + s * G::LANES]) a2 += P[s];
                    if (MISTI_TAB(anc11)[lane + s * G::LANES]) a11 += P[s];
                }
            a2 = g.sum(a2); a11 = g.sum(a11);
            if (do_reset) {
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int r = lane + s * G::LANES;
                    P[s] = r == 2 ? a2 : (r == 11 ? a11 : 0.0);
                }
            }
        }
        double pr = 0.0;
        int src = 0;
        if (has_pulses && act) {
            const double pu0 = pulse_rate(md, params, it, 0), pu1 = pulse_rate(md, params, it, 1);
            pr = pu0 + pu1;
            src = pu0 > 0 ? 0 : 1;
        }
        if (g.any(pr > 0)) {  // a group without a pulse here applies the map with rate 0 = the identity
            const double om = 1.0 - pr;
            const PulseEntry* ent = src == 0 ? MISTI_TAB(pulse0) : MISTI_TAB(pulse1);
            const unsigned char* rp = src == 0 ? MISTI_TAB(pulse0_rowptr) : MISTI_TAB(pulse1_rowptr);
            g.sync();
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                wb[s * G::LANES] = P[s];
            g.sync();
            double pw_om[5], pw_r[5];
            pw_om[0] = 1.0; pw_r[0] = 1.0;
            for (int k = 1; k < 5; ++k) { pw_om[k] = pw_om[k - 1] * om; pw_r[k] = pw_r[k - 1] * pr; }
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                double acc = 0.0;
                if (valid[s]) {
                    const int r = lane + s * G::LANES;
                    for (int e = rp[r]; e < rp[r + 1]; ++e) {
                        const PulseEntry pe = ent[e];
                        double w = (double)pe.mult;
                        for (int k = 0; k < 5; ++k) {
                            if (k == pe.a) w *= pw_om[k];
                            if (k == pe.b) w *= pw_r[k];
                        }
                        acc += w * ysm[pe.col];
                    }
                }
                P[s] = acc;
            }
            g.sync();
        }
    };

    // generator of interval `it` in uniformised form: A = I + M/q with q = max |M_cc| (per group)
    double adiag[RPL], coef[RPL][W], qinv = 1.0, q = 1.0;
    bool mig = false;
    auto set_generator = [&](int it, bool act) {
        double la0 = 1.0, la1 = 1.0, m0 = 0.0, m1 = 0.0;
        if (act) {
            la0 = lc[(PITCH * it) * stride]; la1 = lc[(PITCH * it + 1) * stride];
            if (PITCH == 4) {
                m0 = lc[(PITCH * it + 2) * stride]; m1 = lc[(PITCH * it + 3) * stride];
            } else {
                m0 = band_rate(md, params, it, 0); m1 = band_rate(md, params, it, 1);
            }
            if (!(la0 >= 0.0 && la0 <= DBL_MAX && la1 >= 0.0 && la1 <= DBL_MAX && m0 >= 0.0 && m0 <= DBL_MAX && m1 >= 0.0 &&
                  m1 <= DBL_MAX)) {
                status = MISTI_NONFINITE;
                la0 = la1 = 1.0; m0 = m1 = 0.0;
            }
        }
        mig = m0 + m1 != 0.0;
        double d[RPL], dmax = 0.0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            d[s] = (double)((rc[s] >> 16) & 7u) * la0 + (double)((rc[s] >> 19) & 7u) * la1 +
                   (double)((rc[s] >> 22) & 7u) * m0 + (double)((rc[s] >> 25) & 7u) * m1;
            dmax = d[s] > dmax ? d[s] : dmax;
        }
        q = g.max(dmax);
        if (!(q > 0.0)) q = 1.0;  // no event possible at all: A = I
        qinv = 1.0 / q;
        const double rq[4] = {la0 * qinv, la1 * qinv, m0 * qinv, m1 * qinv};
        const double mine = g.table_entry(rq);
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            adiag[s] = (q - d[s]) * qinv;
#pragma unroll
            for (int e = 0; e < W; ++e) coef[s][e] = g.table_get(mine, (rc[s] >> (4 * e)) & 15u, rq);
        }
    };

    const int n_loop = g.wmax(n_fin > it0 ? n_fin - it0 : 0);
    for (int j = 0; j < n_loop; ++j) {
        const int it = it0 + j;
        bool act = !pending && it < n_fin;  // a group past its own last interval idles on a zero-length interval
        set_generator(it, act);
        double T = act ? times[it] : 0.0;
        if (!(T >= 0.0 && T <= DBL_MAX)) { status = MISTI_NONFINITE; T = 0.0; }
        // rates of 1e5 and more per unit of interval length only come out of a run-away correction; such an interval
        // is not swept here (it would stall the warp for millions of terms): the item is parked for the dense step
        if (q * T > kUnifMaxStiff) {
            if (cont) {
                if (lane == 0) { cont->it = it; cont->nterms = nterms; }
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int r = lane + s * G::LANES;
                    if (r < 48) { cont->P[r] = P[s]; cont->Ia[r] = Ia[s]; cont->Ib[r] = Ib[s]; }
                }
                pending = true;
            } else {
                status = MISTI_STIFF;
            }
            act = false;
            T = 0.0;
        }
        reset_and_pulse(it, act);
        const double qT = q * T;
        const int nsub = g.wmax(qT > kUnifMaxStep ? (int)ceil(qT / kUnifMaxStep) : 1);
        const double lam = nsub == 1 ? qT : qT / nsub;
        const double p0 = exp(-lam), t0 = -expm1(-lam);  // Poisson P(N = 0) and P(N > 0)
        double Iint[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) Iint[s] = 0.0;
        for (int sub = 0; sub < nsub; ++sub) {
            double yk[RPL], P1[RPL];
            g.sync();
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                yk[s] = P[s]; P1[s] = p0 * P[s];
                wb[s * G::LANES] = P[s];
            }
            double p = p0;     // Pois(k; lam)
            double tail = t0;  // P(N > k)
            double r = lam;    // lam / (k + 1): ratio of consecutive Poisson weights
            int k = 0;
            // one term: I += P(N > k-1) y_(k-1);  y_k <- A y_(k-1) (read buffer RO, write buffer WO);  P1 += Pois(k) y_k.
            // Returns true when this group's Poisson tail beyond the term is below kUnifTol.
            auto term = [&](const int RO, const int WO) -> bool {
                g.sync();
                ++k;
                p *= r;
                r = lam * MISTI_RECIP(k + 1);
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    // two short FMA chains per row instead of one long one
                    const double u = fma(coef[s][1], yp[s][1][RO], fma(coef[s][0], yp[s][0][RO], adiag[s] * yk[s]));
                    const double v = fma(coef[s][3], yp[s][3][RO], coef[s][2] * yp[s][2][RO]);
                    const double acc = u + v;
                    Iint[s] = fma(tail, yk[s], Iint[s]);
                    yk[s] = acc;
                    wb[WO + s * G::LANES] = acc;
                    P1[s] = fma(p, acc, P1[s]);
                }
                tail -= p;
                return r < 1.0 && p < kUnifTol * (1.0 - r);
            };
            while (true) {  // the two halves of the ping-pong buffer get compile-time offsets
                if (g.all(term(0, kYStride))) break;
                if (g.all(term(kYStride, 0))) break;
                if (k >= kUnifMaxTerms) { status = MISTI_NONFINITE; break; }
            }
            if (act) nterms += k;
#pragma unroll
            for (int s = 0; s < RPL; ++s) P[s] = P1[s];
        }
        // integralP of this interval joins the running sums; categories 2..6 are muted before the sampling
        // date (:501-506), hence the two accumulators
        if (it < md.sampleDate) {
#pragma unroll
            for (int s = 0; s < RPL; ++s) Ia[s] = fma(Iint[s], qinv, Ia[s]);
        } else {
#pragma unroll
            for (int s = 0; s < RPL; ++s) Ib[s] = fma(Iint[s], qinv, Ib[s]);
        }
    }
    if (g.any(inf_last && !pending)) {
        // no split inside the grid: the last two-population interval is infinite (MigrationInference.py:475-476,
        // 535-538): P1 = 0, integralP = -inv(M) P0 = (1/q) sum_k A^k P0, finite only with migration.
        const int it = numT - 1;
        const bool inf_now = inf_last && !pending;
        reset_and_pulse(it, inf_now);
        set_generator(it, inf_now);
        if (inf_now && !mig && status == MISTI_OK) status = MISTI_INFINITE_COAL_TIME;
        const bool run = inf_now && mig;
        double yk[RPL], Iint[RPL];
        g.sync();
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            yk[s] = run ? P[s] : 0.0;
            Iint[s] = yk[s];
            wb[s * G::LANES] = yk[s];
        }
        double nprev = 0.0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) nprev += yk[s];
        nprev = g.sum(nprev);
        double itot = nprev;
        bool done = !(nprev > 0.0);
        int k = 0, cur = 0;
        while (!g.all(done)) {
            g.sync();
            const int ro = kYStride * cur, wo = kYStride * (cur ^ 1);
            double nk = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                double acc = adiag[s] * yk[s];
#pragma unroll
                for (int e = 0; e < W; ++e) acc += coef[s][e] * yp[s][e][ro];
                yk[s] = acc;
                wb[wo + s * G::LANES] = acc;
                Iint[s] += acc;
                nk += acc;
            }
            cur ^= 1;
            ++k;
            nk = g.sum(nk);
            itot += nk;
            const double rho = nprev > 0.0 ? nk / nprev : 0.0;  // contraction of the remaining mass
            nprev = nk;
            if (!(nk > 0.0) || (rho < 1.0 && nk * rho < kUnifTol * itot * (1.0 - rho))) done = true;
            if (k > 2000000) { if (!done) status = MISTI_NONFINITE; break; }
        }
        if (run) {
            nterms += k;
            const bool pre = it < md.sampleDate;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                if (pre) Ia[s] = fma(Iint[s], qinv, Ia[s]);
                else Ib[s] = fma(Iint[s], qinv, Ib[s]);
                P[s] = 0.0;
            }
        }
    }
    // JAFS = StateToJAF . (sum of the interval integrals) (:501-506)
    double jl[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) jl[c] = 0.0;
#pragma unroll
    for (int s = 0; s < RPL; ++s)
        if (valid[s]) {
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const double w = (double)MISTI_TAB(w44)[c][lane + s * G::LANES];
                jl[c] += w * (c < 2 ? Ia[s] + Ib[s] : Ib[s]);
            }
        }
    const bool post = active && md.splitT < numT;
    if (g.any(post)) {
        const bool do_reset = post && md.splitT == md.sampleDate && md.splitT > 0;  // the reset precedes the collapse (:480-494)
        if (g.any(do_reset)) {
            double a2 = 0.0, a11 = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s]) {
                    if (MISTI_TAB(anc2)[lane + s * G::LANES]) a2 += P[s];
                    if (MISTI_TAB(anc11)[lane + s * G::LANES]) a11 += P[s];
                }
            a2 = g.sum(a2); a11 = g.sum(a11);
            if (do_reset) {
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int r = lane + s * G::LANES;
                    P[s] = r == 2 ? a2 : (r == 11 ? a11 : 0.0);
                }
            }
        }
        // CollapsePops (:518-528): 44 -> 8 block sums
        double P8[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double v = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s] && MISTI_TAB(collapse)[lane + s * G::LANES] == b) v += P[s];
            P8[b] = g.sum(v);
        }
        const double c6 = post ? cpost[0] : 0.0, c3 = post ? cpost[1] : 0.0, c1 = post ? cpost[2] : 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            double a6 = 0.0, a3 = 0.0, a1 = 0.0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                a6 += MISTI_TAB(wg6)[c][b] * P8[b];
                a3 += MISTI_TAB(wg3)[c][b] * P8[b];
                a1 += MISTI_TAB(wg1)[c][b] * P8[b];
            }
            jafs[c] = g.sum(jl[c]) + ((c6 * a6 + c3 * a3) + c1 * a1);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) jafs[c] = g.sum(jl[c]);
    }
    *terms = nterms;
    return pending ? MISTI_STIFF : status;
}

// Normalised spectrum -> log terms used by the composite likelihood (MigrationInference.py:583-613).
// Folded: bins (0+6), (1+5), (2+4), 3; the data vector is folded the same way by the host, so
// logj[4..6] = 0 there.  Returns false if a required log is not finite.
MISTI_HD inline bool jafs_normalise_logs(const double* raw, bool unfolded, double* jn, double* logj) {
    double tot = 0.0;
    for (int c = 0; c < 7; ++c) tot += raw[c];
    for (int c = 0; c < 7; ++c) jn[c] = raw[c] / tot;
    bool ok = true;
    if (unfolded) {
        for (int c = 0; c < 7; ++c) logj[c] = log(jn[c]);
    } else {
        logj[0] = log(jn[0] + jn[6]);
        logj[1] = log(jn[1] + jn[5]);
        logj[2] = log(jn[2] + jn[4]);
        logj[3] = log(jn[3]);
        logj[4] = logj[5] = logj[6] = 0.0;
    }
    for (int c = 0; c < 7; ++c)
        if (!(fabs(logj[c]) <= DBL_MAX)) ok = false;
    return ok;
}

// llh for one data row: const + sum_i d_i log p_i, accumulated in the reference's order (:600-609).
MISTI_HD inline double score_row(const double* drow /* 7 counts (folded by the host if needed) + const */, const double* logj) {
    double llh = drow[7];
    for (int c = 0; c < 7; ++c) llh += drow[c] * logj[c];
    return llh;
}

}  // namespace misti
