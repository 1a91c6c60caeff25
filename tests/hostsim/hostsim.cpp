// tests/hostsim -- TEST-ONLY host build of the __host__ __device__ numerics in
// misti_b200/csrc/misti_math.cuh and misti_model.cuh, so the scalar correction chain (K1's body)
// can be unit-tested against the oracle in a container without a GPU.  It is NOT part of the
// product: libmisti_b200.so has no CPU path and nothing under misti_b200/ loads this library.
#include <climits>
#include <cstring>
#include <vector>
#include "../../misti_b200/csrc/misti_model.cuh"
#include "../../misti_b200/csrc/misti_jsfs.cuh"
#include "../../misti_b200/csrc/misti_optim.cuh"

extern "C" {

void hs_expm3(const double* A, double* E) { misti::mat3_expm(A, E); }

int hs_inv3(const double* A, double* Ainv) { return misti::mat3_inv(A, Ainv) ? 1 : 0; }

// bands: [n][5] = pop(0/1), start, end, value, opt index (-1 fixed); pulses: [n][4] = pop, time, value, opt index
int hs_correct_lambdas(int numT, int splitT, int sampleDate, const double* times, const double* lh, int n_bands,
                       const double* bands, int n_pulses, const double* pulses, int n_params, const double* params,
                       unsigned flags, double mixtureTH, double* lc, double* Pr, int* nfev, int* trace /* [numT][2], nullable */) {
    misti::ModelDesc md;
    std::memset(&md, 0, sizeof(md));
    md.numT = numT; md.splitT = splitT; md.sampleDate = sampleDate;
    md.n_bands = n_bands; md.n_pulses = n_pulses; md.n_params = n_params;
    for (int b = 0; b < n_bands; ++b) {
        md.band_pop[b] = (int)bands[5 * b]; md.band_start[b] = (int)bands[5 * b + 1]; md.band_end[b] = (int)bands[5 * b + 2];
        md.band_val[b] = bands[5 * b + 3]; md.band_opt[b] = (int)bands[5 * b + 4];
    }
    for (int b = 0; b < n_pulses; ++b) {
        md.pulse_pop[b] = (int)pulses[4 * b]; md.pulse_time[b] = (int)pulses[4 * b + 1];
        md.pulse_val[b] = pulses[4 * b + 2]; md.pulse_opt[b] = (int)pulses[4 * b + 3];
    }
    // the grid constants and interval classes the library precomputes on the host (misti_add_grid / misti_add_model)
    std::vector<double> gaux((size_t)numT * misti::kGridAux);
    for (int t = 0; t < numT; ++t) misti::grid_aux_row(lh + 2 * t, t < numT - 1 ? times[t] : 0.0, &gaux[(size_t)t * misti::kGridAux]);
    std::vector<unsigned> cls(numT);
    for (int t = 0; t < numT; ++t) cls[t] = misti::interval_class(md, t);
    if (trace)
        for (int t = 0; t < numT; ++t) { trace[2 * t] = 0; trace[2 * t + 1] = misti::kNoSolve; }
    return misti::correct_lambdas_item<false, false, true>(md, times, lh, params, flags, mixtureTH, lc, 2, 1, Pr, nfev, gaux.data(), nullptr,
                                                           nullptr, cls.data(), nullptr, trace);
}

// CoalescentRates (forward map): lc[numT][2] true rates -> lh_out[numT][2], Pr[(splitT + 1)][3][2]
void hs_coal_rates(int numT, int splitT, const double* times, const double* lc, int n_bands, const double* bands, int n_pulses,
                   const double* pulses, int n_params, const double* params, double mu0, double mu1, double* lh_out, double* Pr);

static void fill_model(misti::ModelDesc& md, int numT, int splitT, int sampleDate, int n_bands, const double* bands,
                       int n_pulses, const double* pulses, int n_params) {
    std::memset(&md, 0, sizeof(md));
    md.numT = numT; md.splitT = splitT; md.sampleDate = sampleDate;
    md.n_bands = n_bands; md.n_pulses = n_pulses; md.n_params = n_params;
    for (int b = 0; b < n_bands; ++b) {
        md.band_pop[b] = (int)bands[5 * b]; md.band_start[b] = (int)bands[5 * b + 1]; md.band_end[b] = (int)bands[5 * b + 2];
        md.band_val[b] = bands[5 * b + 3]; md.band_opt[b] = (int)bands[5 * b + 4];
    }
    for (int b = 0; b < n_pulses; ++b) {
        md.pulse_pop[b] = (int)pulses[4 * b]; md.pulse_time[b] = (int)pulses[4 * b + 1];
        md.pulse_val[b] = pulses[4 * b + 2]; md.pulse_opt[b] = (int)pulses[4 * b + 3];
    }
}

void hs_coal_rates(int numT, int splitT, const double* times, const double* lc, int n_bands, const double* bands, int n_pulses,
                   const double* pulses, int n_params, const double* params, double mu0, double mu1, double* lh_out, double* Pr) {
    misti::ModelDesc md;
    fill_model(md, numT, splitT, 0, n_bands, bands, n_pulses, pulses, n_params);
    std::vector<unsigned> cls(numT);
    for (int t = 0; t < numT; ++t) cls[t] = misti::interval_class(md, t);
    const double mu[2] = {mu0, mu1};
    misti::coalescent_rates_item(md, times, lc, params, mu, cls.data(), lh_out, Pr);
}

static int hs_types_buf[256];
static int* hs_last_types = hs_types_buf;
static int hs_last_nseg = 0;

// segment types of the last hs_jsfs call (1 = swept interval, 2 = closed-form run, 3 = stiff, 4 = infinite)
int hs_segment_types(int* out, int cap) {
    for (int i = 0; i < hs_last_nseg && i < cap; ++i) out[i] = hs_types_buf[i];
    return hs_last_nseg;
}

// expected JSFS (unnormalised raw[7], normalised jn[7]) and llh for one data row, given lc[numT][2]
int hs_jsfs(int numT, int splitT, int sampleDate, const double* times, int n_bands, const double* bands, int n_pulses,
            const double* pulses, int n_params, const double* params, const double* lc, int unfolded,
            const double* drow /* 7 + const */, double* raw, double* jn, double* llh, int* terms) {
    misti::ModelDesc md;
    fill_model(md, numT, splitT, sampleDate, n_bands, bands, n_pulses, pulses, n_params);
    double cpost[3], ysm[misti::kGroupScratch], logj[7];
    misti::post_split_coeffs(md, times, lc, 2, 1, cpost);
    std::vector<double> rec((size_t)(numT + 1) * misti::kRecSlots);
    int nseg = 0;
    std::vector<unsigned> cls(numT);
    for (int t = 0; t < numT; ++t) cls[t] = misti::interval_class(md, t);
    int st = misti::build_segments_item(md, times, params, lc, 2, 1, rec.data(), &nseg, cls.data());
    if (st != MISTI_OK) return st;
    if (hs_last_types) {
        for (int i = 0; i < nseg && i < 256; ++i) hs_last_types[i] = misti::seg_type(misti::seg_meta_bits(rec[i * misti::kRecSlots + 15]));
        hs_last_nseg = nseg;
    }
    misti::SingleLane g;
    static misti::RunTable<misti::SingleLane> runtab;
    runtab.fill(0, 1);
    misti::LaneCtx<misti::SingleLane> L;
    L.init(g, ysm, &runtab);
    st = misti::jsfs_item<misti::SingleLane>(g, L, md, true, params, rec.data(), nseg, cpost, raw, terms);
    if (st != MISTI_OK) return st;
    double jn0;
    if (!misti::jafs_finish(g, ysm, raw, unfolded != 0, &jn0)) return MISTI_NONFINITE;
    for (int c = 0; c < 7; ++c) { jn[c] = ysm[misti::kTailJn + c]; logj[c] = ysm[misti::kTailLog + c]; }
    *llh = misti::score_row(drow, logj);
    return MISTI_OK;
}

// cpfit post-split coefficients from ed = exp(nc1 - nc0): the sequential pass of the correction chain (seq[3], and the
// rates lc[numT][2]) against the lane-group pass over the model's table (grp[3]; one lane here, 16 on the device, and
// `lanes` only shapes the table: the slices are then walked one after the other), and the rate-on-request function
void hs_post_split(int numT, int splitT, const double* times, const double* lh, double ed, int lanes, double* seq, double* grp,
                   double* lc, double* lc_req) {
    misti::ModelDesc md;
    std::memset(&md, 0, sizeof(md));
    md.numT = numT; md.splitT = splitT;
    std::vector<double> gaux((size_t)numT * misti::kGridAux);
    for (int t = 0; t < numT; ++t) misti::grid_aux_row(lh + 2 * t, t < numT - 1 ? times[t] : 0.0, &gaux[(size_t)t * misti::kGridAux]);
    misti::post_split_cpfit_item(md, times, lh, 0.0, log(ed), lc, 2, 1, gaux.data(), seq);
    for (int t = splitT; t < numT; ++t)
        lc_req[t] = misti::post_split_cpfit_rate(md, t, t < numT - 1 ? times[t] : 0.0, &gaux[(size_t)t * misti::kGridAux], lh, ed);
    // the table in the device layout, walked slice by slice: partial sums per slice, survival factors by running product
    const int per = misti::post_split_per(numT, splitT, lanes);
    std::vector<double> tab((size_t)misti::kPostVals * per * lanes + 1);
    misti::post_split_table(numT, splitT, times, gaux.data(), lanes, tab.data());
    misti::SingleLane g;
    double f1 = 1.0;
    grp[0] = grp[1] = grp[2] = 0.0;
    for (int l = 0; l < lanes; ++l) {
        // slice l as a one-lane table: value k of step j at sl[k per + j]
        std::vector<double> sl((size_t)misti::kPostVals * per + 1);
        for (int k = 0; k < misti::kPostVals; ++k)
            for (int j = 0; j < per; ++j) sl[(size_t)k * per + j] = tab[((size_t)k * per + j) * lanes + l];
        // a slice alone = a model whose "infinite interval" has rate 1/0: lh_last = inf keeps the last term out
        double part[3];
        const double lh_inf[2] = {1e300, 1e300};
        misti::post_split_cpfit_group(g, true, sl.data(), per, lh_inf, ed, part);
        double e1 = 1.0;
        for (int j = 0; j < per; ++j) e1 *= (sl[j] + ed * sl[(size_t)per + j]) * (1.0 / (1.0 + ed));
        const double f3 = f1 * f1 * f1;
        grp[0] += f3 * f3 * part[0]; grp[1] += f3 * part[1]; grp[2] += f1 * part[2];
        f1 *= e1;
    }
    const double lam = (1.0 + ed) / (1.0 / lh[2 * (numT - 1)] + ed / lh[2 * (numT - 1) + 1]);
    const double f3 = f1 * f1 * f1;
    grp[0] += f3 * f3 / (6.0 * lam); grp[1] += f3 / (3.0 * lam); grp[2] += f1 / lam;
}

// one simplex of the on-device Nelder-Mead (misti_optim.cuh), driven from the test: propose -> the test evaluates the
// objective -> apply.  maxiter / maxfev < 0 = none.
static misti::NmConfig hs_nm_cfg(int N, double xatol, double fatol, long long maxiter, long long maxfev, int lookahead) {
    misti::NmConfig c;
    c.N = N; c.lookahead = lookahead; c.slots = misti::nm_slots(N, lookahead != 0); c.xatol = xatol; c.fatol = fatol;
    c.maxiter = maxiter < 0 ? LLONG_MAX : maxiter; c.maxfev = maxfev < 0 ? LLONG_MAX : maxfev;
    return c;
}

int hs_nm_slots(int N, int lookahead) { return misti::nm_slots(N, lookahead != 0); }

int hs_nm_propose(int N, double xatol, double fatol, long long maxiter, long long maxfev, int lookahead, double* sim, double* fsim,
                  long long* iters, long long* fcalls, int* status, int* phase, double* pts) {
    return misti::nm_propose(hs_nm_cfg(N, xatol, fatol, maxiter, maxfev, lookahead), sim, fsim, iters, fcalls, status, phase, pts);
}

void hs_nm_apply(int N, double xatol, double fatol, long long maxiter, long long maxfev, int lookahead, double* sim, double* fsim,
                 long long* iters, long long* fcalls, int* status, int* phase, const double* pts, const double* fv) {
    misti::nm_apply(hs_nm_cfg(N, xatol, fatol, maxiter, maxfev, lookahead), sim, fsim, iters, fcalls, status, phase, pts, fv, false);
}

// numpy's Generator(PCG64) continued from `state` (4 words, see misti_fit_opts.rng_state): n doubles of uniform(low, high)
void hs_pcg64_uniform(unsigned long long* state, double low, double high, int n, double* out) {
    misti::Pcg64 r = {state[0], state[1], state[2], state[3]};
    for (int i = 0; i < n; ++i) out[i] = misti::pcg64_uniform(r, low, high);
    state[0] = r.s_hi; state[1] = r.s_lo; state[2] = r.inc_hi; state[3] = r.inc_lo;
}

// one basin-hopping walker (misti_optim.cuh: bh_advance) around the Nelder-Mead step logic, driven from the test:
// the walker's state lives in the arrays the test owns; returns 1 while the walker goes on
int hs_bh_advance(int N, int niter, int interval, double T, double target, double factor, double* x, double* best_x, double* scal /* energy, best_f, step */,
                  int* flags /* ok, best_ok, done */, long long* cnt /* nfev, failures, nstep, naccept, hop */, unsigned long long* rng,
                  double* sim, double* fsim, long long* iters, long long* fcalls, int* status, int* phase) {
    misti::BhConfig c;
    c.niter = niter; c.interval = interval; c.beta = T != 0 ? 1.0 / T : misti::kInf; c.target = target; c.factor = factor; c.stepsize0 = scal[2];
    misti::Pcg64 r = {rng[0], rng[1], rng[2], rng[3]};
    misti::BhWalker w;
    w.x = x; w.best_x = best_x; w.energy = scal; w.best_f = scal + 1; w.step = scal + 2;
    w.ok = flags; w.best_ok = flags + 1; w.done = flags + 2;
    w.nfev = cnt; w.failures = cnt + 1; w.nstep = cnt + 2; w.naccept = cnt + 3; w.hop = cnt + 4; w.rng = &r;
    const bool go = misti::bh_advance(c, N, w, sim, fsim, iters, fcalls, status, phase);
    rng[0] = r.s_hi; rng[1] = r.s_lo; rng[2] = r.inc_hi; rng[3] = r.inc_lo;
    return go ? 1 : 0;
}

}  // extern "C"
