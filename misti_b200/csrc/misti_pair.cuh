// misti_pair.cuh -- expected joint SFS of one evaluation item by a PAIR of lanes (device only; large batches).
//
// Same reference path and the same algorithm as misti_jsfs.cuh (uniformisation sweeps for intervals with migration,
// closed-form projector sums for runs without, pulses, ancient-sample reset, collapse + closed-form post-split tail,
// MigrationInference.py:467-540, TwoPopulations.py:188-262, 336-377), mapped differently: the chain is symmetric under the
// exchange of the two demes, so lane 0 of a pair owns one state of every two-cycle of that symmetry (plus the two fixed
// states) and lane 1 the images, both run the same straight-line code (tools/gen_pair_tables.py -> misti_pair_code.h) and
// differ only in the order in which they read the rate table.  The state vector, the running integrals and the generator
// live in registers; the partner's values come by shuffle; shared memory is only touched by pulses.  The 16-lane kernel of
// misti_jsfs.cuh is bound by the shared-memory pipe (13 wavefronts per mat-vec and item); this one is bound by FP64 issue.
//
// What it does not do: stiff segments and the infinite last interval -- an item that meets one is handed to the 16-lane
// kernel untouched (redo list).
#pragma once
#include "misti_jsfs.cuh"
#include "misti_pair_code.h"

namespace misti {

constexpr int kPairN = MISTI_PAIR_N;
static __device__ const unsigned char d_pair_row[2][kPairN] = MISTI_PAIR_ROW_INIT;
static __device__ const unsigned char d_pair_fixed[kPairN] = MISTI_PAIR_FIXED_INIT;

struct PairResult {
    double raw[7];  // unnormalised spectrum (both lanes hold all entries)
    int nterms;
    bool redo;      // met a segment this kernel leaves to the 16-lane one
};

// All lanes of the warp call this together, every pair with its own item (`active` = false: the pair runs along idle).
// `ysm`: 48 doubles of shared memory per pair (pulses only).
__device__ __forceinline__ void jsfs_pair_item(const ModelDesc& md, bool active, const double* __restrict__ params,
                                               const double* __restrict__ rec, int nseg, const double* cpost, double* ysm,
                                               PairResult* res) {
    constexpr unsigned FULL = 0xffffffffu;
    const int role = threadIdx.x & 1;
    double y[kPairN], Ia[kPairN];
#pragma unroll
    for (int i = 0; i < kPairN; ++i) {
        y[i] = (i == MISTI_PAIR_START_ROW && role == MISTI_PAIR_START_ROLE) ? 1.0 : 0.0;  // both genome-1 lineages in deme 0, genome-2 in deme 1 (:469-471)
        Ia[i] = 0.0;
    }
    // occupancy integrals: categories 2..6 are muted before the sampling date (:501-506).  One accumulator: when the
    // sweep crosses the sampling date, what has been summed so far is folded into the two categories that count there.
    double jpre0 = 0.0, jpre1 = 0.0;
    bool in_pre = false;
    int nterms = 0;
    bool redo = false;
    int nown = active ? nseg : 0;

    // AncientSampleP0 (TwoPopulations.py:246-262)
    auto ancient_reset = [&](bool do_reset) {
        double a2 = 0.0, a11 = 0.0;
#pragma unroll
        for (int i = 0; i < kPairN; ++i) {
            const int r = d_pair_row[role][i];
            const bool mine = !(role && d_pair_fixed[i]);
            if (mine && MISTI_TAB(anc2)[r]) a2 += y[i];
            if (mine && MISTI_TAB(anc11)[r]) a11 += y[i];
        }
        a2 += __shfl_xor_sync(FULL, a2, 1);
        a11 += __shfl_xor_sync(FULL, a11, 1);
        if (do_reset) {
#pragma unroll
            for (int i = 0; i < kPairN; ++i) {
                const int r = d_pair_row[role][i];
                y[i] = r == 2 ? a2 : (r == 11 ? a11 : 0.0);
            }
        }
    };
    // PulseMigration (:361-377) before interval `it`, through the pair's shared scratch (rare: once per pulse and item)
    auto pulse = [&](int it, bool do_pulse) {
        double pr = 0.0;
        int src = 0;
        if (do_pulse) {
            const double pu0 = pulse_rate(md, params, it, 0), pu1 = pulse_rate(md, params, it, 1);
            pr = pu0 + pu1;
            src = pu0 > 0 ? 0 : 1;
        }
        const double om = 1.0 - pr;
        const PulseEntry* ent = src == 0 ? MISTI_TAB(pulse0) : MISTI_TAB(pulse1);
        const unsigned char* rp = src == 0 ? MISTI_TAB(pulse0_rowptr) : MISTI_TAB(pulse1_rowptr);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kPairN; ++i)
            if (!(role && d_pair_fixed[i])) ysm[d_pair_row[role][i]] = y[i];
        __syncwarp();
        double pw_om[5], pw_r[5];
        pw_om[0] = 1.0; pw_r[0] = 1.0;
        for (int k = 1; k < 5; ++k) { pw_om[k] = pw_om[k - 1] * om; pw_r[k] = pw_r[k - 1] * pr; }
#pragma unroll
        for (int i = 0; i < kPairN; ++i) {
            const int r = d_pair_row[role][i];
            double acc = 0.0;
            for (int e = rp[r]; e < rp[r + 1]; ++e) {
                const PulseEntry pe = ent[e];
                double w = (double)pe.mult;
                for (int k = 0; k < 5; ++k) {
                    if (k == pe.a) w *= pw_om[k];
                    if (k == pe.b) w *= pw_r[k];
                }
                acc += w * ysm[pe.col];
            }
            y[i] = acc;
        }
        __syncwarp();
    };

    const int n_loop = __reduce_max_sync(FULL, nown);
    double meta_next = nown > 0 ? rec[15] : 0.0;  // the meta word of the next record is fetched a segment ahead
    for (int sg = 0; sg < n_loop; ++sg) {
        const bool have = sg < nown;
        const double* r = rec + (long)sg * kRecSlots;
        const unsigned long long meta = have ? seg_meta_bits(meta_next) : 0ull;
        meta_next = sg + 1 < nown ? r[kRecSlots + 15] : 0.0;
        int type = seg_type(meta);
        const int it = seg_it(meta);
        if (type == SEG_STIFF || type == SEG_INF) {  // not for this kernel: the whole item goes to the 16-lane one
            redo = true;
            nown = 0;
            type = SEG_NOP;
        }
        const bool do_reset = type != SEG_NOP && (meta & kSegReset) != 0, do_pulse = type != SEG_NOP && (meta & kSegPulse) != 0;
        if (__any_sync(FULL, do_reset)) ancient_reset(do_reset);
        if (__any_sync(FULL, do_pulse)) pulse(it, do_pulse);  // a pair without a pulse applies the map with rate 0 = the identity
        if (type != SEG_NOP) {
            const bool pre = (meta & kSegPre) != 0;
            if (in_pre && !pre) {  // crossing the sampling date
                double jw[7];
                double X[kPairN];
#pragma unroll
                for (int i = 0; i < kPairN; ++i) X[i] = Ia[i];
                MISTI_PAIR_ZERO_FIXED(X, role == 0);
                MISTI_PAIR_TAIL_W(X, jw);
                jpre0 += jw[0]; jpre1 += jw[1];
#pragma unroll
                for (int i = 0; i < kPairN; ++i) Ia[i] = 0.0;
            }
            in_pre = pre;
        }
        const unsigned m_mig = __ballot_sync(FULL, type == SEG_MIG);
        if (type == SEG_MIG) {
            // generator in uniformised form, A = I + M/q: lane 1 reads the coefficient table with the rate kinds swapped
            double cf[10], dg[MISTI_PAIR_NDIAG];
#pragma unroll
            for (int c = 0; c < 10; ++c) cf[c] = r[c ^ role];
            MISTI_PAIR_DIAG(dg, cf);
            const double qinv = r[10], lam_own = r[11], p0_own = r[13], t0_own = r[14];
            const int K_own = seg_K(meta), nsub_own = seg_nsub(meta);
            const int nsub = __reduce_max_sync(m_mig, nsub_own);
            for (int sub = 0; sub < nsub; ++sub) {
                const bool live = sub < nsub_own;
                double lam = live ? lam_own : 0.0;
                double p = live ? p0_own : 1.0;  // Pois(k; lam)
                double tail = live ? t0_own : 0.0;  // P(N > k)
                const int Ks = live ? K_own : 0;
                const int Kmax = __reduce_max_sync(m_mig, Ks);
                double P1[kPairN];
#pragma unroll
                for (int i = 0; i < kPairN; ++i) P1[i] = p * y[i];
                double rr = lam;  // lam / (k + 1): ratio of consecutive Poisson weights
                // one term: I += P(N > k-1) y_(k-1) / q;  y_k <- A y_(k-1);  P1 += Pois(k) y_k.  A pair whose own series
                // has ended goes on with zero weights (exact no-op on P1 and the integrals)
#pragma unroll 1  // (two terms per trip save the 46 register moves of y <- A y but spill: 0.193 -> 0.222 ms)
                for (int k = 1; k <= Kmax; ++k) {
                    if (k > Ks) { lam = 0.0; rr = 0.0; tail = 0.0; }
                    const double tq = tail * qinv;
                    p *= rr;
                    rr = lam * MISTI_RECIP(k + 1);
                    MISTI_PAIR_TERM(y, Ia, P1, cf, dg, tq, p, m_mig);
                    tail -= p;
                }
                if (live) nterms += Ks;
#pragma unroll
                for (int i = 0; i < kPairN; ++i) y[i] = P1[i];
            }
        }
        const unsigned m_run = __ballot_sync(FULL, type == SEG_RUN);
        if (type == SEG_RUN) {
            // a run of intervals without migration: P <- sum_ab e_ab G_ab P, integral += sum_ab c_ab G_ab P; under the
            // exchange of the demes G_ab becomes G_ba, so lane 1 reads the coefficients with a and b swapped
            const unsigned char swp[8] = MISTI_PAIR_ABSWAP_INIT;
            double C[8], E[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) C[a] = r[role ? swp[a] : a];
            E[0] = 1.0;
#pragma unroll
            for (int a = 1; a < 8; ++a) E[a] = r[7 + (role ? swp[a] : a)];
            MISTI_PAIR_RUN(y, Ia, C, E, m_run);
            nterms += 1;
        }
    }

    // JAFS = StateToJAF . (sum of the interval integrals) (:501-506) + the one-population tail after the split: with
    // P8 = CollapsePops(P) (:518-528) that is sum_b V[c][b] P8[b], V = c6 WG6 + c3 WG3 + c1 WG1
    const bool post = active && !redo && md.splitT < md.numT;
    {
        const bool do_reset = post && md.splitT == md.sampleDate && md.splitT > 0;  // the reset precedes the collapse (:480-494)
        if (__any_sync(FULL, do_reset)) ancient_reset(do_reset);
    }
    MISTI_PAIR_ZERO_FIXED(Ia, role == 0);
    MISTI_PAIR_ZERO_FIXED(y, role == 0);
    double jl[7];
    MISTI_PAIR_TAIL_W(Ia, jl);
    if (in_pre) {
#pragma unroll
        for (int c = 2; c < 7; ++c) jl[c] = 0.0;
    }
    jl[0] += jpre0; jl[1] += jpre1;
    if (post) {
        const double c6 = cpost[0], c3 = cpost[1], c1 = cpost[2];
        double V[7][8];
        MISTI_PAIR_V(V, c6, c3, c1);
        MISTI_PAIR_TAIL_V(V, y, jl);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) res->raw[c] = jl[c] + __shfl_xor_sync(FULL, jl[c], 1);
    res->nterms = nterms;
    res->redo = redo;
}

}  // namespace misti
