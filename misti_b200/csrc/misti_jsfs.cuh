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
struct SingleLane {  // test-only host build: one lane owns all 44 rows
    static constexpr int LANES = 1;
    MISTI_HD int lane() const { return 0; }
    MISTI_HD void sync() const {}
    MISTI_HD double max(double v) const { return v; }
    MISTI_HD double sum(double v) const { return v; }
};

#if defined(__CUDACC__)
struct WarpLanes {  // one warp per item: lane l owns rows l and l + 32
    static constexpr int LANES = 32;
    __device__ int lane() const { return threadIdx.x & 31; }
    __device__ void sync() const { __syncwarp(); }
    __device__ double max(double v) const {
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ double sum(double v) const {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};
#endif

constexpr double kUnifMaxStep = 32.0;           // largest q*T handled in one uniformisation sweep
constexpr double kUnifTol = 1.3877787807814457e-17;  // 2^-56: truncation of the Poisson tail

// Post-split coefficients (run by ONE thread; lc addressed like in correct_lambdas_item):
//   cpost[k] = sum_{i>=splitT} exp(-a_k x_i) (1 - exp(-a_k lam_i T_i)) / (a_k lam_i),  x_i = sum_{j<i} lam_j T_j,
// with the last interval infinite (MigrationInference.py:530-540: P1 = 0 there), a = (6, 3, 1).
MISTI_HD inline void post_split_coeffs(const ModelDesc& md, const double* times, const double* lc, long stride, double* cpost) {
    double c6 = 0, c3 = 0, c1 = 0;
    double e1 = 1.0;  // exp(-x)
    for (int t = md.splitT; t < md.numT; ++t) {
        const double lam = lc[(2 * t) * stride];
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

// Expected JSFS of one item.  All lanes of the group call this together; `ysm` is a scratch area of
// 2*44 doubles shared by the group.  On return every lane holds the UNNORMALISED spectrum in
// jafs[0..6] (MigrationInference.JAFSpectrum's return value) and the number of mat-vecs in *terms.
template <class G>
MISTI_D inline int jsfs_item(const G& g, const ModelDesc& md, const double* times, const double* params, const double* lc,
                              long stride, const double* cpost, double* ysm, double* jafs, int* terms) {
    constexpr int RPL = (44 + G::LANES - 1) / G::LANES;
    const int lane = g.lane();
    int row[RPL];
    bool valid[RPL];
    unsigned code[RPL][MISTI_ELL_WIDTH];  // col | kind << 8 | cnt << 16
    unsigned dcode[RPL];                  // diagonal multiplicities, 8 bits per rate kind
    unsigned wcode[RPL];                  // W44 column, 2 bits per SFS category
    double P[RPL];
    double jl[7];
    for (int c = 0; c < 7; ++c) jl[c] = 0.0;
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int r = lane + s * G::LANES;
        valid[s] = r < 44;
        row[s] = valid[s] ? r : 0;
        dcode[s] = 0; wcode[s] = 0;
        for (int e = 0; e < MISTI_ELL_WIDTH; ++e) {
            const EllEntry en = MISTI_TAB(ell)[row[s]][e];
            code[s][e] = valid[s] ? (en.col | (en.kind << 8) | (en.cnt << 16)) : 0u;
        }
        if (valid[s]) {
            for (int k = 0; k < 4; ++k) dcode[s] |= (unsigned)MISTI_TAB(diag)[r][k] << (8 * k);
            for (int c = 0; c < 7; ++c) wcode[s] |= (unsigned)MISTI_TAB(w44)[c][r] << (2 * c);
        }
        P[s] = (valid[s] && r == 2) ? 1.0 : 0.0;  // both genome-1 lineages in deme 0, genome-2 in deme 1 (:469-471)
    }
    int nterms = 0;
    int status = MISTI_OK;
    const int numT = md.numT;
    const int n2 = md.splitT < numT ? md.splitT : numT;  // number of two-population intervals
    for (int it = 0; it < n2; ++it) {
        if (it == md.sampleDate && it > 0) {  // AncientSampleP0 (TwoPopulations.py:246-262); identity on the start vector
            double a2 = 0.0, a11 = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s]) {
                    if (MISTI_TAB(anc2)[row[s]]) a2 += P[s];
                    if (MISTI_TAB(anc11)[row[s]]) a11 += P[s];
                }
            a2 = g.sum(a2); a11 = g.sum(a11);
#pragma unroll
            for (int s = 0; s < RPL; ++s) P[s] = !valid[s] ? 0.0 : (row[s] == 2 ? a2 : (row[s] == 11 ? a11 : 0.0));
        }
        const double pu0 = pulse_rate(md, params, it, 0), pu1 = pulse_rate(md, params, it, 1);
        if (pu0 + pu1 > 0) {  // PulseMigration (TwoPopulations.py:361-377)
            const double r = pu0 + pu1, om = 1.0 - r;
            const int src = pu0 > 0 ? 0 : 1;
            const PulseEntry* ent = src == 0 ? MISTI_TAB(pulse0) : MISTI_TAB(pulse1);
            const unsigned char* rp = src == 0 ? MISTI_TAB(pulse0_rowptr) : MISTI_TAB(pulse1_rowptr);
            g.sync();
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s]) ysm[row[s]] = P[s];
            g.sync();
            double pw_om[5], pw_r[5];
            pw_om[0] = 1.0; pw_r[0] = 1.0;
            for (int k = 1; k < 5; ++k) { pw_om[k] = pw_om[k - 1] * om; pw_r[k] = pw_r[k - 1] * r; }
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                double acc = 0.0;
                if (valid[s])
                    for (int e = rp[row[s]]; e < rp[row[s] + 1]; ++e) {
                        const PulseEntry pe = ent[e];
                        acc += (double)pe.mult * pw_om[pe.a] * pw_r[pe.b] * ysm[pe.col];
                    }
                P[s] = acc;
            }
            g.sync();
        }
        const double la0 = lc[(2 * it) * stride], la1 = lc[(2 * it + 1) * stride];
        const double m0 = band_rate(md, params, it, 0), m1 = band_rate(md, params, it, 1);
        const double rate[4] = {la0, la1, m0, m1};
        double d[RPL], dmax = 0.0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            d[s] = (double)(dcode[s] & 255u) * la0 + (double)((dcode[s] >> 8) & 255u) * la1 +
                   (double)((dcode[s] >> 16) & 255u) * m0 + (double)((dcode[s] >> 24) & 255u) * m1;
            dmax = d[s] > dmax ? d[s] : dmax;
        }
        const double q = g.max(dmax);
        if (!(q > 0.0) || !(q <= DBL_MAX)) { status = MISTI_NONFINITE; break; }
        const double qinv = 1.0 / q;
        double adiag[RPL], coef[RPL][MISTI_ELL_WIDTH];
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            adiag[s] = (q - d[s]) * qinv;
#pragma unroll
            for (int e = 0; e < MISTI_ELL_WIDTH; ++e)
                coef[s][e] = (double)(code[s][e] >> 16) * rate[(code[s][e] >> 8) & 3u] * qinv;
        }
        double Iacc[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) Iacc[s] = 0.0;
        const bool last = it == numT - 1;
        if (!last) {
            const double qT = q * times[it];
            int nsub = 1;
            if (qT > kUnifMaxStep) nsub = (int)ceil(qT / kUnifMaxStep);
            const double lam = qT / nsub;
            const double p0 = exp(-lam);
            for (int sub = 0; sub < nsub; ++sub) {
                double yk[RPL], S[RPL], P1[RPL], I[RPL];
                int cur = 0;
                g.sync();
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    yk[s] = P[s]; S[s] = 0.0; P1[s] = p0 * P[s]; I[s] = 0.0;
                    if (valid[s]) ysm[row[s]] = P[s];
                }
                double p = p0;
                int k = 0;
                while (true) {
                    g.sync();
                    const double* yr = ysm + 44 * cur;
                    double* yw = ysm + 44 * (cur ^ 1);
                    ++k;
                    p *= lam / k;
#pragma unroll
                    for (int s = 0; s < RPL; ++s) {
                        double acc = adiag[s] * yk[s];
#pragma unroll
                        for (int e = 0; e < MISTI_ELL_WIDTH; ++e) acc += coef[s][e] * yr[code[s][e] & 255u];
                        S[s] += yk[s];
                        yk[s] = acc;
                        if (valid[s]) yw[row[s]] = acc;
                        P1[s] += p * acc;
                        I[s] += p * S[s];
                    }
                    cur ^= 1;
                    if (k + 1 > lam && p * (k + 1) < kUnifTol * (k + 1 - lam)) break;
                    if (k > 4096) { status = MISTI_NONFINITE; break; }
                }
                nterms += k;
#pragma unroll
                for (int s = 0; s < RPL; ++s) { P[s] = P1[s]; Iacc[s] += I[s]; }
                if (status != MISTI_OK) break;
            }
#pragma unroll
            for (int s = 0; s < RPL; ++s) Iacc[s] *= qinv;
        } else {
            // infinite last interval before any split (MigrationInference.py:475-476, 535-538):
            // integralP = -inv(M) P0 = (1/q) sum_k A^k P0; needs migration to be finite.
            if (m0 + m1 == 0.0) { status = MISTI_INFINITE_COAL_TIME; break; }
            double yk[RPL];
            int cur = 0;
            g.sync();
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                yk[s] = P[s];
                Iacc[s] = P[s];
                if (valid[s]) ysm[row[s]] = P[s];
            }
            double nprev = 0.0, itot = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s) nprev += yk[s];
            nprev = g.sum(nprev);
            itot = nprev;
            int k = 0;
            while (nprev > 0.0) {
                g.sync();
                const double* yr = ysm + 44 * cur;
                double* yw = ysm + 44 * (cur ^ 1);
                double nk = 0.0;
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    double acc = adiag[s] * yk[s];
#pragma unroll
                    for (int e = 0; e < MISTI_ELL_WIDTH; ++e) acc += coef[s][e] * yr[code[s][e] & 255u];
                    yk[s] = acc;
                    if (valid[s]) yw[row[s]] = acc;
                    Iacc[s] += acc;
                    nk += acc;
                }
                cur ^= 1;
                ++k;
                nk = g.sum(nk);
                itot += nk;
                const double rho = nk / nprev;  // contraction of the remaining mass
                nprev = nk;
                if (rho < 1.0 && nk * rho < kUnifTol * itot * (1.0 - rho)) break;
                if (k > 2000000) { status = MISTI_NONFINITE; break; }
            }
            nterms += k;
#pragma unroll
            for (int s = 0; s < RPL; ++s) { Iacc[s] *= qinv; P[s] = 0.0; }
            if (status != MISTI_OK) break;
        }
        // JAFS += StateToJAF . integralP; categories 2..6 are muted before the sampling date (:501-506)
        const int cmax = it < md.sampleDate ? 2 : 7;
#pragma unroll
        for (int s = 0; s < RPL; ++s)
#pragma unroll
            for (int c = 0; c < 7; ++c)
                if (c < cmax) jl[c] += (double)((wcode[s] >> (2 * c)) & 3u) * Iacc[s];
    }
    if (status == MISTI_OK && md.splitT < numT) {
        if (md.splitT == md.sampleDate && md.splitT > 0) {  // the reset precedes the collapse (:480-494)
            double a2 = 0.0, a11 = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s]) {
                    if (MISTI_TAB(anc2)[row[s]]) a2 += P[s];
                    if (MISTI_TAB(anc11)[row[s]]) a11 += P[s];
                }
            a2 = g.sum(a2); a11 = g.sum(a11);
#pragma unroll
            for (int s = 0; s < RPL; ++s) P[s] = !valid[s] ? 0.0 : (row[s] == 2 ? a2 : (row[s] == 11 ? a11 : 0.0));
        }
        // CollapsePops (:518-528): 44 -> 8 block sums
        double P8[8];
        for (int b = 0; b < 8; ++b) {
            double v = 0.0;
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (valid[s] && MISTI_TAB(collapse)[row[s]] == b) v += P[s];
            P8[b] = g.sum(v);
        }
        const double c6 = cpost[0], c3 = cpost[1], c1 = cpost[2];
        for (int c = 0; c < 7; ++c) {
            double a6 = 0.0, a3 = 0.0, a1 = 0.0;
            for (int b = 0; b < 8; ++b) {
                a6 += MISTI_TAB(wg6)[c][b] * P8[b];
                a3 += MISTI_TAB(wg3)[c][b] * P8[b];
                a1 += MISTI_TAB(wg1)[c][b] * P8[b];
            }
            jafs[c] = g.sum(jl[c]) + ((c6 * a6 + c3 * a3) + c1 * a1);
        }
    } else {
        for (int c = 0; c < 7; ++c) jafs[c] = g.sum(jl[c]);
    }
    *terms = nterms;
    return status;
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
