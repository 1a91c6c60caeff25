// misti_model.cuh -- model descriptor shared by host and device, parameter mapping, and the
// per-item coalescence-rate correction chain (the body of kernel K1).
//
// Reference: MigrationInference.SetModel / MapParameters (MigrationInference.py:229-298),
// CorrectLambdas (:305-378), SmoothConst (:387-405).
#pragma once
#include <stddef.h>
#include "misti_math.cuh"

#include "../../include/misti_b200.h"  // flags, status codes and capacity limits are part of the C ABI

namespace misti {

struct ModelDesc {
    int numT;        // number of intervals (rate pairs); times has numT-1 entries
    int splitT;      // first one-population interval
    int sampleDate;  // interval index at which genome 2 was sampled
    int n_bands;
    int n_pulses;
    int n_params;
    int grid_off;    // offset of this model's grid inside the pooled arrays (in intervals)
    int cls_off;     // offset of this model's interval classes inside the pooled class array (see interval_class)
    int post_off;    // offset (in doubles) of this model's post-split table inside the pooled table array (post_split_table)
    int post_per;    // post-split intervals per lane of a 16-lane group in that table
    int band_pop[MISTI_MAX_BANDS], band_start[MISTI_MAX_BANDS], band_end[MISTI_MAX_BANDS], band_opt[MISTI_MAX_BANDS];
    int pulse_pop[MISTI_MAX_PULSES], pulse_time[MISTI_MAX_PULSES], pulse_opt[MISTI_MAX_PULSES];
    double band_val[MISTI_MAX_BANDS];
    double pulse_val[MISTI_MAX_PULSES];
};

// mi[t][pop] after MapParameters: optimised bands take params[opt], fixed bands their value.
MISTI_HD inline double band_rate(const ModelDesc& md, const double* params, int t, int pop) {
    double v = 0.0;
    for (int b = 0; b < md.n_bands; ++b)
        if (md.band_pop[b] == pop && t >= md.band_start[b] && t < md.band_end[b])
            v = md.band_opt[b] >= 0 ? params[md.band_opt[b]] : md.band_val[b];
    return v;
}

MISTI_HD inline double pulse_rate(const ModelDesc& md, const double* params, int t, int pop) {
    double v = 0.0;
    for (int b = 0; b < md.n_pulses; ++b)
        if (md.pulse_pop[b] == pop && md.pulse_time[b] == t)
            v = md.pulse_opt[b] >= 0 ? params[md.pulse_opt[b]] : md.pulse_val[b];
    return v;
}

// Interval class of a model: which band feeds deme 0 / deme 1 and which pulse leaves deme 0 / deme 1 at interval t, as
// four bytes (index + 1, 0 = none) -- the band / pulse loops above evaluated once per model on the host instead of
// four times per interval and item on the device.
MISTI_HD inline unsigned interval_class(const ModelDesc& md, int t) {
    unsigned w = 0;
    for (int b = 0; b < md.n_bands; ++b)
        if (t >= md.band_start[b] && t < md.band_end[b]) {
            const int sh = md.band_pop[b] == 0 ? 0 : 8;
            w = (w & ~(255u << sh)) | ((unsigned)(b + 1) << sh);
        }
    for (int b = 0; b < md.n_pulses; ++b)
        if (md.pulse_time[b] == t) {
            const int sh = md.pulse_pop[b] == 0 ? 16 : 24;
            w = (w & ~(255u << sh)) | ((unsigned)(b + 1) << sh);
        }
    return w;
}

// mi[t][0..1] and pu[t][0..1] of interval t; cls (nullable) = the model's interval classes
MISTI_HD inline void interval_rates(const ModelDesc& md, const unsigned* cls, const double* params, int t, double* mi, double* pu) {
    if (!cls) {
        mi[0] = band_rate(md, params, t, 0); mi[1] = band_rate(md, params, t, 1);
        pu[0] = pulse_rate(md, params, t, 0); pu[1] = pulse_rate(md, params, t, 1);
        return;
    }
    const unsigned w = cls[t];
    mi[0] = mi[1] = pu[0] = pu[1] = 0.0;
    if (w == 0) return;
    for (int k = 0; k < 2; ++k) {
        const unsigned b = (w >> (8 * k)) & 255u;
        if (b) mi[k] = md.band_opt[b - 1] >= 0 ? params[md.band_opt[b - 1]] : md.band_val[b - 1];
        const unsigned q = (w >> (16 + 8 * k)) & 255u;
        if (q) pu[k] = md.pulse_opt[q - 1] >= 0 ? params[md.pulse_opt[q - 1]] : md.pulse_val[q - 1];
    }
}

// Post-split rates in cpfit mode (MigrationInference.py:356-374) and, with them, the post-split closed-form coefficients
// cpost[3] (see post_split_coeffs in misti_jsfs.cuh).  nc0, nc1 = the reference's running "log probabilities" at the split.
// pnc_t = (exp(-T lh0) + exp((nc1 - nc0) - T lh1)) / (1 + exp(nc1 - nc0)), lam_t = -log(pnc_t) / T, and nc0, nc1 both drop
// by T lam_t -- so d = nc1 - nc0 never changes and the intervals are independent of each other:
// pnc_t = (E0_t + e^d E1_t) / (1 + e^d) with the grid constants E_g = exp(-T lh_g).  The last rate
// (pr0 + pr1) / (pr0 / lh0 + pr1 / lh1), pr_k = exp(nc_k), is (1 + e^d) / (1 / lh0 + e^d / lh1).
// One post-split rate of that pass, from ed = exp(nc1 - nc0) (`ga` = grid_aux_row of interval t, T its length)
MISTI_HD inline double post_split_cpfit_rate(const ModelDesc& md, int t, double T, const double* ga, const double* lh, double ed) {
    if (t == md.numT - 1) return (1.0 + ed) / (1.0 / lh[2 * t] + ed / lh[2 * t + 1]);
    if (T == 0) return 1.0;
    const double u = (ga[0] + ed * ga[1]) * (1.0 / (1.0 + ed));
    return -log(u) * ga[4];
}

MISTI_HD inline void post_split_cpfit_item(const ModelDesc& md, const double* times, const double* lh, double nc0, double nc1,
                                           double* lc, int pitch, long stride, const double* gaux, double* cpost) {
    const int numT = md.numT, splitT = md.splitT;
    if (!(splitT < numT)) return;
    const double ed = exp(nc1 - nc0), wn = 1.0 / (1.0 + ed);
    double c6 = 0, c3 = 0, c1 = 0, e1 = 1.0;  // e1 = exp(-sum of lam T so far)
    for (int t = splitT; t < numT - 1; ++t) {
        const double T = times[t];
        if (T == 0) { lc[(pitch * t) * stride] = 1; lc[(pitch * t + 1) * stride] = 1; continue; }
        double ga[kGridAux];
        if (gaux) { for (int i = 0; i < kGridAux; ++i) ga[i] = gaux[kGridAux * t + i]; }
        else grid_aux_row(lh + 2 * t, T, ga);
        const double u = (ga[0] + ed * ga[1]) * wn;   // pnc = exp(-lam T)
        const double q1 = (ga[2] + ed * ga[3]) * wn;  // 1 - pnc, free of cancellation
        const double z = -log(u);
        const double lam = z * ga[4];
        lc[(pitch * t) * stride] = lam;
        lc[(pitch * t + 1) * stride] = lam;
        const double il = z > 0 ? T / z : 0.0;  // 1 / lam
        const double e3 = e1 * e1 * e1;
        const double q3 = q1 * (1.0 + u + u * u), q6 = q3 * (1.0 + u * u * u);  // 1 - u^3, 1 - u^6
        c1 += z > 0 ? e1 * q1 * il : e1 * T;
        c3 += z > 0 ? e3 * q3 * (il * (1.0 / 3.0)) : e3 * T;
        c6 += z > 0 ? (e3 * e3) * q6 * (il * (1.0 / 6.0)) : (e3 * e3) * T;
        e1 *= u;
    }
    {
        const int t = numT - 1;
        const double lam = (1.0 + ed) / (1.0 / lh[2 * t] + ed / lh[2 * t + 1]);
        lc[(pitch * t) * stride] = lam;
        lc[(pitch * t + 1) * stride] = lam;
        const double il = 1.0 / lam, e3 = e1 * e1 * e1;
        c1 += e1 * il; c3 += e3 * (il * (1.0 / 3.0)); c6 += (e3 * e3) * (il * (1.0 / 6.0));
    }
    if (cpost) { cpost[0] = c6; cpost[1] = c3; cpost[2] = c1; }
}

// The coefficients of that pass alone, from ed = exp(nc1 - nc0) (the rates are not needed on the path; the gather kernel
// computes them on request): the same arithmetic in the same order as post_split_cpfit_item, so the two agree bit for bit.
// `gaux` = the grid's aux rows (grid_aux_row).  One thread; the intervals are independent of each other given ed, so the
// loop pipelines (four logs in flight).
MISTI_HD inline void post_split_cpfit_coeffs(const ModelDesc& md, const double* times, const double* lh, double ed,
                                             const double* gaux, double* cpost) {
    const int numT = md.numT, splitT = md.splitT;
    if (!(splitT < numT)) return;
    const double wn = 1.0 / (1.0 + ed);
    double c6 = 0, c3 = 0, c1 = 0, e1 = 1.0;
#pragma unroll 4
    for (int t = splitT; t < numT - 1; ++t) {
        const double T = times[t];
        if (T == 0) continue;
        const double* ga = gaux + kGridAux * t;
        const double u = (ga[0] + ed * ga[1]) * wn;
        const double q1 = (ga[2] + ed * ga[3]) * wn;
        const double z = -log(u);
        const double il = z > 0 ? T / z : 0.0;
        const double e3 = e1 * e1 * e1;
        const double q3 = q1 * (1.0 + u + u * u), q6 = q3 * (1.0 + u * u * u);
        c1 += z > 0 ? e1 * q1 * il : e1 * T;
        c3 += z > 0 ? e3 * q3 * (il * (1.0 / 3.0)) : e3 * T;
        c6 += z > 0 ? (e3 * e3) * q6 * (il * (1.0 / 6.0)) : (e3 * e3) * T;
        e1 *= u;
    }
    {
        const int t = numT - 1;
        const double lam = (1.0 + ed) / (1.0 / lh[2 * t] + ed / lh[2 * t + 1]);
        const double il = 1.0 / lam, e3 = e1 * e1 * e1;
        c1 += e1 * il; c3 += e3 * (il * (1.0 / 3.0)); c6 += (e3 * e3) * (il * (1.0 / 6.0));
    }
    cpost[0] = c6; cpost[1] = c3; cpost[2] = c1;
}

// The finite post-split intervals of a model as the lane groups of the JSFS kernel read them (post_split_cpfit_group in
// misti_jsfs.cuh): lane l of `lanes` owns the `per` consecutive intervals splitT + l per + j, j < per, and value k
// (E0, E1, 1 - E0, 1 - E1 of grid_aux_row, and T) of its j-th interval sits at out[(k per + j) lanes + l] -- one coalesced
// load per value and step.  Slots past the last finite interval and zero-length intervals (skipped by the reference,
// MigrationInference.py:358-360) hold the neutral entry E = 1, 1 - E = 0, T = 0.
constexpr int kPostVals = 5;
MISTI_HD inline int post_split_per(int numT, int splitT, int lanes) {
    const int n = splitT < numT ? numT - 1 - splitT : 0;
    return (n + lanes - 1) / lanes;
}
inline void post_split_table(int numT, int splitT, const double* times, const double* gaux, int lanes, double* out) {
    const int per = post_split_per(numT, splitT, lanes);
    for (int l = 0; l < lanes; ++l)
        for (int j = 0; j < per; ++j) {
            const int t = splitT + l * per + j;
            const bool have = t < numT - 1 && times[t] != 0;
            const double v[kPostVals] = {have ? gaux[kGridAux * t] : 1.0, have ? gaux[kGridAux * t + 1] : 1.0,
                                         have ? gaux[kGridAux * t + 2] : 0.0, have ? gaux[kGridAux * t + 3] : 0.0,
                                         have ? times[t] : 0.0};
            for (int k = 0; k < kPostVals; ++k) out[((size_t)k * per + j) * lanes + l] = v[k];
        }
}

// Checkpoint of an interrupted correction chain (one per item, global memory).  The chain is strictly sequential and, where
// the reference's own solver runs away (rates ~1e7: trust-region solves that spend their whole budget of 200 evaluations on
// matrices that need 20 squarings), one item can take ten times as long as its neighbours; inside the on-device optimiser
// such an item would hold up every simplex of the round.  There the chain may YIELD at an interval boundary once it has run
// for its time slice: everything the next interval needs is the state below (the rates written so far stay where they are),
// and the next round resumes it -- same instructions on the same data, so the result does not depend on where, or whether,
// a chain was cut.
struct ChainCkpt {
    int active;        // 1 = interrupted: resume at interval t_next
    int t_next, nfev;
    int sr_k[2], sr_done[2];
    double P0[6], sr_lam[2], sr_nc[2], sr_time[2];
};
struct ChainResume {
    ChainCkpt* ck;
    long long t_start, budget_ns;  // device clock at the start of this slice; <= 0: never yield
};
#define MISTI_PENDING 7  // internal (never leaves the optimiser): the item's chain was interrupted and resumes next round

MISTI_HD inline bool chain_should_yield(const ChainResume* rs, bool coop) {
#if defined(__CUDA_ARCH__)
    if (!rs || rs->budget_ns <= 0) return false;
    long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    bool y = now - rs->t_start > rs->budget_ns;
    if (coop) y = __shfl_sync(0xFu << (threadIdx.x & 28u), (int)y, threadIdx.x & 28u) != 0;  // the four lanes of an item agree
    return y;
#else
    (void)rs; (void)coop;
    return false;
#endif
}

// lc is addressed as lc[(pitch*t+g)*stride] (pitch >= 2 values per interval); times[numT-1]; lh[numT][2].
// Pr (nullable): [splitT+1][3][2] trajectory of the 3-state chains (MigrationInference.py:309,350).
// gaux (nullable): [numT][kGridAux] per-interval constants of the grid (grid_aux_row).
// cpost (nullable): in cpfit mode the post-split closed-form coefficients (see post_split_coeffs in misti_jsfs.cuh)
// fall out of the post-split pass for free (exp(-lam T) is the fitted non-coalescence probability itself);
// *cpost_done tells the caller whether they were written.
// nc_out (nullable): in cpfit mode do NOT run the post-split pass here but return nc0, nc1 (and *cpost_done = true): the
// caller has the pass done elsewhere (post_split_cpfit_group in the JSFS kernel; the post-split rates themselves are
// then only computed on request, post_split_cpfit_rate).
// trace (nullable): [numT][2] per interval the evaluations (scipy's nfev) and the termination status of its least-squares
// solve, kNoSolve where the interval has a closed form -- the iterate-level record the golden vectors of
// tests/golden/solver.json pin (CorrectLambda.py:85, 260, 303, 305).
// COOP (device only): four lanes run the item together, see eval_fj in misti_math.cuh.
template <bool COOP = false, bool RESUME = false, bool TRACE = false>
MISTI_HD inline int correct_lambdas_item(const ModelDesc& md, const double* times, const double* lh, const double* params,
                                         unsigned flags, double mixtureTH, double* lc, int pitch, long stride, double* Pr,
                                         int* nfev_out, const double* gaux = nullptr, double* cpost = nullptr,
                                         bool* cpost_done = nullptr, const unsigned* cls = nullptr, double* nc_out = nullptr,
                                         int* trace = nullptr, const ChainResume* rs = nullptr, int* align = nullptr,
                                         int align_total = 0, int align_every = 1) {
    if (cpost_done) *cpost_done = false;
    const bool correct = flags & MISTI_FLAG_CORRECT, cpfit = flags & MISTI_FLAG_CPFIT;
    int nfev = 0;
    for (int i = 0; i < md.n_params; ++i)
        if (params[i] < 0) return MISTI_NEGATIVE_PARAM;
    IntervalState st;
    st.P0[0][0] = 1; st.P0[0][1] = 0; st.P0[0][2] = 0;
    st.P0[1][0] = 0; st.P0[1][1] = 1; st.P0[1][2] = 0;
    if (Pr) { Pr[0] = 1; Pr[1] = 0; Pr[2] = 0; Pr[3] = 1; Pr[4] = 0; Pr[5] = 0; }
    double nc0 = 0, nc1 = 0;
    const int numT = md.numT, splitT = md.splitT;
    // SmoothConst (:380-405) as a streaming pass: per genome the current run of equal PSMC rates [sk, t), the rate that
    // defines it, and the sums of lc T and T over it; a run is averaged and written back when it ends.
    // (one set of scalars per genome and a step function called with a constant genome index: nothing is subscripted at
    // run time, so the state stays in registers)
    const bool smooth = (flags & MISTI_FLAG_SMOOTH) != 0;
    struct SmoothRun { int k; bool done; double lam, nc, time; };
    SmoothRun sr0 = {0, !smooth, lh[0], 0.0, 0.0}, sr1 = {0, !smooth, lh[1], 0.0, 0.0};
    // (RESUME / TRACE are compile-time switches: the plain batched evaluation carries neither the checkpoint code nor the trace)
    if (!TRACE) trace = nullptr;
    if (!RESUME) rs = nullptr;
    int t_first = 0;
    if (RESUME && rs && rs->ck->active) {  // resume an interrupted chain (see ChainCkpt)
        const ChainCkpt& k = *rs->ck;
        t_first = k.t_next;
        nfev = k.nfev;
        for (int i = 0; i < 3; ++i) { st.P0[0][i] = k.P0[i]; st.P0[1][i] = k.P0[3 + i]; }
        sr0.k = k.sr_k[0]; sr0.done = k.sr_done[0] != 0; sr0.lam = k.sr_lam[0]; sr0.nc = k.sr_nc[0]; sr0.time = k.sr_time[0];
        sr1.k = k.sr_k[1]; sr1.done = k.sr_done[1] != 0; sr1.lam = k.sr_lam[1]; sr1.nc = k.sr_nc[1]; sr1.time = k.sr_time[1];
        nc0 = (st.P0[0][0] + st.P0[0][1]) + st.P0[0][2];
        nc1 = (st.P0[1][0] + st.P0[1][1]) + st.P0[1][2];
    }
    auto smooth_step = [&](SmoothRun& r, int g, int t, double lcv) {
        if (r.done) return;
        const double lhv = lh[2 * t + g];
        if (!(fabs(lhv - r.lam) < 1e-10 && t < numT - 1)) {  // the run [k, t) ends here
            if (t > r.k) {
                const double avg = r.nc / r.time;
                for (int i = r.k; i < t; ++i) lc[(pitch * i + g) * stride] = avg;
            }
            if (t >= numT - 1) { r.done = true; return; }  // the last interval is never smoothed (the reference loops forever here)
            r.k = t; r.lam = lhv; r.nc = 0.0; r.time = 0.0;
        }
        r.nc += lcv * times[t];
        r.time += times[t];
    };
    for (int t = t_first; t < splitT; ++t) {
#if defined(__CUDA_ARCH__)
        // `align` (nullable; one-block-per-SM launches of the plain kernel): the threads of the block enter every interval
        // together, so that the SM's warps share the instruction stream.  An unaligned barrier (threads of a warp may arrive
        // from different places); the caller makes up the arrivals of a chain that ends early (*align counts them).
        if (!COOP && !RESUME && !TRACE && align && t % align_every == 0) {
            asm volatile("barrier.sync 1;" ::: "memory");
            ++*align;
        }
#endif
        if (RESUME && t > t_first && chain_should_yield(rs, COOP)) {  // time slice used up: park the chain at this interval boundary
            ChainCkpt& k = *rs->ck;
            bool writer = true;
#if defined(__CUDA_ARCH__)
            if (COOP) writer = (threadIdx.x & 3u) == 0;
#endif
            if (writer) {
                k.t_next = t; k.nfev = nfev;
                for (int i = 0; i < 3; ++i) { k.P0[i] = st.P0[0][i]; k.P0[3 + i] = st.P0[1][i]; }
                k.sr_k[0] = sr0.k; k.sr_done[0] = sr0.done; k.sr_lam[0] = sr0.lam; k.sr_nc[0] = sr0.nc; k.sr_time[0] = sr0.time;
                k.sr_k[1] = sr1.k; k.sr_done[1] = sr1.done; k.sr_lam[1] = sr1.lam; k.sr_nc[1] = sr1.nc; k.sr_time[1] = sr1.time;
                k.active = 1;
            }
            *nfev_out = nfev;
            return MISTI_PENDING;
        }
        double mi_t[2], pu_t[2];
        interval_rates(md, cls, params, t, mi_t, pu_t);
        const double pu0 = pu_t[0], pu1 = pu_t[1];
        const double pu = pu0 + pu1;
        if (pu > 0) {  // closed-form pulse on the 3-state chains (:315-323)
            const int a = pu0 > 0 ? 0 : 1, b = 1 - a;
            for (int k = 0; k < 2; ++k) {
                double* p = st.P0[k];
                const double om = 1 - pu;
                const double qa = p[a] * (om * om);
                const double qb = (p[a] * (pu * pu) + p[b]) + p[2] * pu;
                const double q2 = ((p[a] * 2) * om) * pu + p[2] * om;
                p[a] = qa; p[b] = qb; p[2] = q2;
            }
        }
        double l[2] = {lh[2 * t], lh[2 * t + 1]};
        if (correct) {
            st.lh[0] = lh[2 * t]; st.lh[1] = lh[2 * t + 1];
            st.T = times[t];
            st.mu[0] = mi_t[0]; st.mu[1] = mi_t[1];
            const int nfev_before = nfev;
            int solve_status = kNoSolve;
            const bool ok = solve_interval<COOP>(&st, cpfit, mixtureTH, l, &nfev, gaux ? gaux + kGridAux * t : nullptr,
                                                 trace ? &solve_status : nullptr);
            if (trace) { trace[2 * t] = nfev - nfev_before; trace[2 * t + 1] = solve_status; }
            if (!ok) {
                lc[(pitch * t) * stride] = l[0];
                lc[(pitch * t + 1) * stride] = l[1];
                *nfev_out = nfev;
                return MISTI_CORRECTION_FAILED;
            }
        }
        lc[(pitch * t) * stride] = l[0];
        lc[(pitch * t + 1) * stride] = l[1];
        smooth_step(sr0, 0, t, l[0]);
        smooth_step(sr1, 1, t, l[1]);
        if (Pr) {
            double* q = Pr + 6 * (t + 1);
            for (int s = 0; s < 3; ++s) { q[2 * s] = st.P0[0][s]; q[2 * s + 1] = st.P0[1][s]; }
        }
        nc0 = (st.P0[0][0] + st.P0[0][1]) + st.P0[0][2];  // reference quirk: a probability used as a log (:353-354)
        nc1 = (st.P0[1][0] + st.P0[1][1]) + st.P0[1][2];
    }
#if defined(__CUDA_ARCH__)
    // a chain shorter than the block's longest one makes up its arrivals HERE, before its post-split work: the threads with
    // longer chains would otherwise wait at their next barrier until this one has finished everything else
    if (!COOP && !RESUME && !TRACE && align)
        for (; *align < align_total; ++*align) asm volatile("barrier.sync 1;" ::: "memory");
#endif
    if (!sr0.done && splitT > sr0.k) {
        const double avg = sr0.nc / sr0.time;
        for (int i = sr0.k; i < splitT; ++i) lc[(pitch * i) * stride] = avg;
    }
    if (!sr1.done && splitT > sr1.k) {
        const double avg = sr1.nc / sr1.time;
        for (int i = sr1.k; i < splitT; ++i) lc[(pitch * i + 1) * stride] = avg;
    }
    if (cpfit && splitT < numT) {
        if (nc_out) {  // the caller has the post-split pass run elsewhere
            nc_out[0] = nc0; nc_out[1] = nc1;
            if (cpost_done) *cpost_done = true;
        } else {
            post_split_cpfit_item(md, times, lh, nc0, nc1, lc, pitch, stride, gaux, cpost);
            if (cpost && cpost_done) *cpost_done = true;
        }
    } else {
        for (int t = splitT; t < numT - 1; ++t) {
            const double T = times[t];
            if (T == 0) { lc[(pitch * t) * stride] = 1; lc[(pitch * t + 1) * stride] = 1; continue; }
            double lam;
            const int nfev_before = nfev;
            int fit_status = kNoSolve;
            const bool ok = fit_single_pop<COOP>(lh + 2 * t, T, nc0, nc1, &lam, &nfev, &fit_status);
            if (trace) { trace[2 * t] = nfev - nfev_before; trace[2 * t + 1] = fit_status; }
            if (!ok) { *nfev_out = nfev; return MISTI_NONFINITE; }
            lc[(pitch * t) * stride] = lam;
            lc[(pitch * t + 1) * stride] = lam;
            nc0 += -T * lam;
            nc1 += -T * lam;
        }
        {
            const int t = numT - 1;
            const double pr0 = exp(nc0), pr1 = exp(nc1);
            const double lam = (pr0 + pr1) / (pr0 / lh[2 * t] + pr1 / lh[2 * t + 1]);
            lc[(pitch * t) * stride] = lam;
            lc[(pitch * t + 1) * stride] = lam;
        }
    }
    *nfev_out = nfev;
    return MISTI_OK;
}

// The forward map MigrationInference.CoalescentRates (MigrationInference.py:542-564 -> CorrectLambda.CoalRates,
// CorrectLambda.py:112-122): `lc` = the true model rates [numT][2]; out: lh_out[numT][2] = the rates PSMC would see
// (only the intervals before the split change) and Pr[(splitT + 1)][3][2], the trajectory of the two 3-state chains.
// mu[2]: the reference never sets the migration rates of its CorrectLambda helper in this method, so EVERY interval is
// propagated with the rates its last CorrectLambdas call left there (those of interval splitT - 1) -- the caller passes them.
MISTI_HD inline void coalescent_rates_item(const ModelDesc& md, const double* times, const double* lc, const double* params,
                                           const double* mu, const unsigned* cls, double* lh_out, double* Pr) {
    const int numT = md.numT, splitT = md.splitT;
    for (int i = 0; i < 2 * numT; ++i) lh_out[i] = lc[i];
    double P0[2][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}};
    for (int t = 0; t < splitT; ++t) {
        double mi_t[2], pu_t[2];
        interval_rates(md, cls, params, t, mi_t, pu_t);
        const double pu0 = pu_t[0], pu1 = pu_t[1], pu = pu0 + pu1;
        if (pu > 0) {  // closed-form pulse on the 3-state chains (:549-556)
            const int a = pu0 > 0 ? 0 : 1, b = 1 - a;
            for (int k = 0; k < 2; ++k) {
                double* p = P0[k];
                const double om = 1 - pu;
                const double qa = p[a] * (om * om);
                const double qb = (p[a] * (pu * pu) + p[b]) + p[2] * pu;
                const double q2 = ((p[a] * 2) * om) * pu + p[2] * om;
                p[a] = qa; p[b] = qb; p[2] = q2;
            }
        }
        if (t == 0 && Pr)
            for (int s = 0; s < 3; ++s) { Pr[2 * s] = P0[0][s]; Pr[2 * s + 1] = P0[1][s]; }
        const double T = times[t];
        double M[9], E[9];
        corr_matrix(lc + 2 * t, mu, T, M);
        mat3_expm(M, E);
        for (int k = 0; k < 2; ++k) {
            double p[3];
            mat3_vec(E, P0[k], p);
            const double before = (P0[k][0] + P0[k][1]) + P0[k][2], after = (p[0] + p[1]) + p[2];
            lh_out[2 * t + k] = -log(after / before) / T;
            P0[k][0] = p[0]; P0[k][1] = p[1]; P0[k][2] = p[2];
        }
        if (Pr) {
            double* q = Pr + 6 * (t + 1);
            for (int s = 0; s < 3; ++s) { q[2 * s] = P0[0][s]; q[2 * s + 1] = P0[1][s]; }
        }
    }
}

}  // namespace misti
