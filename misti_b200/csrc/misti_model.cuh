// misti_model.cuh -- model descriptor shared by host and device, parameter mapping, and the
// per-item coalescence-rate correction chain (the body of kernel K1).
//
// Reference: MigrationInference.SetModel / MapParameters (MigrationInference.py:229-298),
// CorrectLambdas (:305-378), SmoothConst (:387-405).
#pragma once
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
    int pad_;
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

// lc is addressed as lc[(pitch*t+g)*stride] (pitch >= 2 values per interval); times[numT-1]; lh[numT][2].
// Pr (nullable): [splitT+1][3][2] trajectory of the 3-state chains (MigrationInference.py:309,350).
// gaux (nullable): [numT][kGridAux] per-interval constants of the grid (grid_aux_row).
// cpost (nullable): in cpfit mode the post-split closed-form coefficients (see post_split_coeffs in misti_jsfs.cuh)
// fall out of the post-split pass for free (exp(-lam T) is the fitted non-coalescence probability itself);
// *cpost_done tells the caller whether they were written.
MISTI_HD inline int correct_lambdas_item(const ModelDesc& md, const double* times, const double* lh, const double* params,
                                         unsigned flags, double mixtureTH, double* lc, int pitch, long stride, double* Pr,
                                         int* nfev_out, const double* gaux = nullptr, double* cpost = nullptr,
                                         bool* cpost_done = nullptr) {
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
    for (int t = 0; t < splitT; ++t) {
        const double pu0 = pulse_rate(md, params, t, 0), pu1 = pulse_rate(md, params, t, 1);
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
        if (!correct) {
            lc[(pitch * t) * stride] = lh[2 * t];
            lc[(pitch * t + 1) * stride] = lh[2 * t + 1];
        } else {
            st.lh[0] = lh[2 * t]; st.lh[1] = lh[2 * t + 1];
            st.T = times[t];
            st.mu[0] = band_rate(md, params, t, 0); st.mu[1] = band_rate(md, params, t, 1);
            double l[2];
            const bool ok = solve_interval(&st, cpfit, mixtureTH, l, &nfev, gaux ? gaux + kGridAux * t : nullptr);
            lc[(pitch * t) * stride] = l[0];
            lc[(pitch * t + 1) * stride] = l[1];
            if (!ok) { *nfev_out = nfev; return MISTI_CORRECTION_FAILED; }
        }
        if (Pr) {
            double* q = Pr + 6 * (t + 1);
            for (int s = 0; s < 3; ++s) { q[2 * s] = st.P0[0][s]; q[2 * s + 1] = st.P0[1][s]; }
        }
        nc0 = (st.P0[0][0] + st.P0[0][1]) + st.P0[0][2];  // reference quirk: a probability used as a log (:353-354)
        nc1 = (st.P0[1][0] + st.P0[1][1]) + st.P0[1][2];
    }
    if (cpfit && splitT < numT) {
        // Post-split rates, cpfit mode (:356-374): pnc_t = (exp(-T lh0) + exp((nc1 - nc0) - T lh1)) / (1 + exp(nc1 - nc0)),
        // lam_t = -log(pnc_t) / T, and nc0, nc1 both drop by T lam_t -- so d = nc1 - nc0 never changes and the intervals are
        // independent of each other: pnc_t = (E0_t + e^d E1_t) / (1 + e^d) with the grid constants E_g = exp(-T lh_g).
        // The last rate (pr0 + pr1) / (pr0 / lh0 + pr1 / lh1), pr_k = exp(nc_k), is (1 + e^d) / (1 / lh0 + e^d / lh1).
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
        if (cpost) {
            cpost[0] = c6; cpost[1] = c3; cpost[2] = c1;
            if (cpost_done) *cpost_done = true;
        }
    } else {
        for (int t = splitT; t < numT - 1; ++t) {
            const double T = times[t];
            if (T == 0) { lc[(pitch * t) * stride] = 1; lc[(pitch * t + 1) * stride] = 1; continue; }
            double lam;
            if (!fit_single_pop(lh + 2 * t, T, nc0, nc1, &lam, &nfev)) { *nfev_out = nfev; return MISTI_NONFINITE; }
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
    if (flags & MISTI_FLAG_SMOOTH) {  // SmoothConst for both genomes (:380-405)
        for (int g = 0; g < 2; ++g) {
            int k = 0;
            double lam = lh[g];
            double time = 0.0, nc = 0.0;
            while (k < splitT) {
                int j = k;
                while (fabs(lh[2 * j + g] - lam) < 1e-10 && j < numT - 1) {
                    nc += lc[(pitch * j + g) * stride] * times[j];
                    time += times[j];
                    ++j;
                    if (j == splitT) break;
                }
                if (j == k) break;  // splitT == numT: the reference loops forever here; we stop
                const double avg = nc / time;
                for (int i = k; i < j; ++i) lc[(pitch * i + g) * stride] = avg;
                lam = lh[2 * j + g];
                nc = 0.0; time = 0.0;
                k = j;
            }
        }
    }
    *nfev_out = nfev;
    return MISTI_OK;
}

}  // namespace misti
