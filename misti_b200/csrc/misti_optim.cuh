// misti_optim.cuh -- Nelder-Mead on the device: the decision logic of one simplex, one step at a time.
//
// The reference fits by scipy.optimize.minimize(method='Nelder-Mead') around the objective, one evaluation per call
// (MigrationInference.Solve, MigrationInference.py:718-733).  misti_b200/optim.py advances many simplices in lock step
// from the host, one batched evaluation per step; here the same step is taken ON the device so that a fit is a stream
// of kernel launches (propose -> evaluate -> apply) without a host round trip per step.  Every simplex takes exactly
// the decisions of scipy 1.18.1 (_optimize.py:_minimize_neldermead: rho = 1, chi = 2, psi = 0.5, sigma = 0.5, initial
// simplex x0 (1.05) or 0.00025, strict / non-strict comparisons, stable re-ordering, termination
// max|x_i - x_0| <= xatol and max|f_i - f_0| <= fatol, maxiter / maxfev) with the arithmetic of numpy (no fused
// multiply-adds: every product and sum is rounded on its own), so the iterates, the iteration count and scipy's
// evaluation count are those of the host drivers, bit for bit, given the same objective values.
//
// Like the host driver with speculative=True, a step evaluates the four candidates (expansion, reflection, outside and
// inside contraction) together; a shrink takes one more round for its N points.  scipy enforces maxfev INSIDE an
// iteration (its objective wrapper raises once the budget is spent, _optimize.py:549-559, and the iteration is abandoned
// where it stands): the second evaluation of a step, or the rest of a shrink, is then dropped -- reproduced here.
// Host and device compile this header (tests/hostsim drives it on the CPU against scipy itself).
#pragma once
#include "misti_math.cuh"

namespace misti {

struct NmConfig {
    int N;             // parameters per simplex
    int slots;         // points a simplex may submit per round: nm_slots(N, lookahead)
    int lookahead;     // two iterations per round (N <= kNmLookaheadMaxN)
    double xatol, fatol;
    long long maxiter, maxfev;  // LLONG_MAX = none
};

enum { NM_INIT = 0, NM_STEP = 1, NM_SHRINK = 2, NM_DONE = 3 };

// rounded-once arithmetic (numpy evaluates a * b - c * d as two products and a difference)
#if defined(__CUDA_ARCH__)
MISTI_HD inline double nm_mul(double a, double b) { return __dmul_rn(a, b); }
MISTI_HD inline double nm_add(double a, double b) { return __dadd_rn(a, b); }
MISTI_HD inline double nm_sub(double a, double b) { return __dsub_rn(a, b); }
MISTI_HD inline double nm_div(double a, double b) { return __ddiv_rn(a, b); }
#else
MISTI_HD inline double nm_mul(double a, double b) { volatile double r = a * b; return r; }
MISTI_HD inline double nm_add(double a, double b) { volatile double r = a + b; return r; }
MISTI_HD inline double nm_sub(double a, double b) { volatile double r = a - b; return r; }
MISTI_HD inline double nm_div(double a, double b) { volatile double r = a / b; return r; }
#endif

// stable sort of the N + 1 vertices by objective value (numpy argsort(kind="stable")): insertion sort
MISTI_HD inline void nm_sort(int N, double* sim, double* fsim) {
    for (int i = 1; i <= N; ++i) {
        const double fi = fsim[i];
        double row[MISTI_MAX_PARAMS];
        for (int k = 0; k < N; ++k) row[k] = sim[i * N + k];
        int j = i - 1;
        while (j >= 0 && fsim[j] > fi) {
            fsim[j + 1] = fsim[j];
            for (int k = 0; k < N; ++k) sim[(j + 1) * N + k] = sim[j * N + k];
            --j;
        }
        fsim[j + 1] = fi;
        for (int k = 0; k < N; ++k) sim[(j + 1) * N + k] = row[k];
    }
}

// Termination tests at the top of scipy's loop (_optimize.py:833-836 and the while condition); returns true when the
// simplex stops and sets *status (0 converged, 1 maxfev, 2 maxiter).  numpy's max propagates NaN (inf - inf), and a NaN
// compares false: not converged.
MISTI_HD inline bool nm_retire(const NmConfig& c, const double* sim, const double* fsim, long long iters, long long fcalls, int* status) {
    const int N = c.N;
    const bool budget = fcalls < c.maxfev && iters < c.maxiter;
    double dx = 0.0, df = 0.0;
    bool nan_x = false, nan_f = false;
    for (int j = 1; j <= N; ++j) {
        for (int k = 0; k < N; ++k) {
            const double d = fabs(sim[j * N + k] - sim[k]);
            if (d != d) nan_x = true;
            dx = d > dx ? d : dx;
        }
        const double d = fabs(fsim[0] - fsim[j]);
        if (d != d) nan_f = true;
        df = d > df ? d : df;
    }
    const bool conv = !nan_x && !nan_f && dx <= c.xatol && df <= c.fatol;
    if (budget && !conv) return false;
    *status = (budget && conv) ? 0 : (fcalls >= c.maxfev ? 1 : 2);
    return true;
}

// reflection-type candidates of a sorted simplex (_optimize.py:846-874), in the order of the decision codes: expansion,
// reflection, outside and inside contraction -> pts[4][N]
MISTI_HD inline void nm_candidates(int N, const double* sim, double* pts) {
    for (int k = 0; k < N; ++k) {
        double xbar = sim[k];
        for (int j = 1; j < N; ++j) xbar = nm_add(xbar, sim[j * N + k]);  // np.add.reduce(sim[:-1], 0): row by row
        xbar = nm_div(xbar, (double)N);
        const double last = sim[N * N + k];
        pts[0 * N + k] = nm_sub(nm_mul(3.0, xbar), nm_mul(2.0, last));   // (1 + rho chi) xbar - rho chi last
        pts[1 * N + k] = nm_sub(nm_mul(2.0, xbar), nm_mul(1.0, last));   // (1 + rho) xbar - rho last
        pts[2 * N + k] = nm_sub(nm_mul(1.5, xbar), nm_mul(0.5, last));   // (1 + psi rho) xbar - psi rho last
        pts[3 * N + k] = nm_add(nm_mul(0.5, xbar), nm_mul(0.5, last));   // (1 - psi) xbar + psi last
    }
}

// Look-ahead (as in misti_b200/optim.py): with the candidates of the current step a simplex also submits the candidates
// of the NEXT step for every way the current one can end -- accepted candidate o, landing at rank k of the sorted
// simplex: expansion (rank 0), reflection (ranks 0..N-1), either contraction (ranks 0..N): 3N + 3 scenarios of 4 points
// -- so that one round of launches advances it by two iterations.  Decisions and counts are unchanged.
constexpr int kNmLookaheadMaxN = 4;  // 4 (3N + 4) points per simplex and round: 28 ... 64
MISTI_HD inline int nm_scenarios(int N) { return 3 * N + 3; }
MISTI_HD inline int nm_scenario(int N, int o, int k) { return (o == 0 ? 0 : (o == 1 ? 1 : (o == 2 ? 1 + N : 2 + 2 * N))) + k; }
MISTI_HD inline int nm_slots(int N, bool lookahead) { return lookahead ? 4 * (1 + nm_scenarios(N)) : (N + 1 > 4 ? N + 1 : 4); }

// The points a simplex wants evaluated in this round, written to pts[slots][N]; returns how many (0 = none: done).
// sim[(N+1)][N], fsim[N+1] are sorted except in phase NM_INIT (sim[0] = x0) and NM_SHRINK (vertices 1..N just moved).
MISTI_HD inline int nm_propose(const NmConfig& c, double* sim, const double* fsim, const long long* iters, const long long* fcalls,
                               int* status, int* phase, double* pts) {
    const int N = c.N;
    if (*phase == NM_DONE) return 0;
    if (*phase == NM_INIT) {  // scipy's default simplex (_optimize.py:775-801)
        for (int j = 1; j <= N; ++j)
            for (int k = 0; k < N; ++k) {
                double y = sim[k];
                if (k == j - 1) y = y != 0 ? nm_mul(1 + 0.05, y) : 0.00025;
                sim[j * N + k] = y;
            }
        for (int i = 0; i < (N + 1) * N; ++i) pts[i] = sim[i];
        return c.maxfev < N + 1 ? (int)c.maxfev : N + 1;
    }
    if (*phase == NM_SHRINK) {  // the vertices the budget still covers (nm_apply moved no more than that, plus one)
        const long long left = c.maxfev - *fcalls;
        const int n = left < N ? (int)left : N;
        for (int j = 1; j <= n; ++j)
            for (int k = 0; k < N; ++k) pts[(j - 1) * N + k] = sim[j * N + k];
        return n;
    }
    if (nm_retire(c, sim, fsim, *iters, *fcalls, status)) {
        *phase = NM_DONE;
        return 0;
    }
    nm_candidates(N, sim, pts);
    if (!c.lookahead) return 4;
    double t[(kNmLookaheadMaxN + 1) * kNmLookaheadMaxN];
    for (int o = 0; o < 4; ++o) {
        const int kmax = o == 0 ? 0 : (o == 1 ? N - 1 : N);
        for (int k = 0; k <= kmax; ++k) {  // the simplex after candidate o has replaced the worst vertex and sorted to rank k
            for (int j = 0; j <= N; ++j) {
                const double* src = j < k ? sim + j * N : (j == k ? pts + o * N : sim + (j - 1) * N);
                for (int i = 0; i < N; ++i) t[j * N + i] = src[i];
            }
            nm_candidates(N, t, pts + (long)(4 + 4 * nm_scenario(N, o, k)) * N);
        }
    }
    return 4 * (1 + nm_scenarios(N));
}

MISTI_HD inline double nm_clean(double v, bool negate) {  // NaN counts as +inf; the device hands over llh = -objective
    v = negate ? -v : v;
    return v != v ? kInf : v;
}

// One reflection-type step (scipy's decision tree, _optimize.py:846-896) with the objective values fv[4] of pts[4][N].
// Returns the accepted candidate (0..3; *rank = where it landed in the sorted simplex), -1 after a shrink move (phase
// NM_SHRINK unless the budget is spent), -2 when the budget cut the iteration short.
MISTI_HD inline int nm_step(const NmConfig& c, double* sim, double* fsim, long long* iters, long long* fcalls, int* phase,
                            const double* pts, const double* fv_in, bool negate, int* rank) {
    const int N = c.N;
    const long long left = c.maxfev - *fcalls;  // evaluations the budget still covers
    const double fxe = nm_clean(fv_in[0], negate), fxr = nm_clean(fv_in[1], negate), fxc = nm_clean(fv_in[2], negate),
                 fxcc = nm_clean(fv_in[3], negate);
    const bool better = fxr < fsim[0];
    const bool second = fxr < fsim[N - 1];
    const bool take_e = better && fxe < fxr;
    const bool take_r = (better && !(fxe < fxr)) || (!better && second);
    const bool contract = !better && !second;
    const bool outside = contract && fxr < fsim[N], inside = contract && !(fxr < fsim[N]);
    const bool take_c = outside && fxc <= fxr, take_cc = inside && fxcc < fsim[N];
    const int which = take_e ? 0 : (take_r ? 1 : (take_c ? 2 : (take_cc ? 3 : -1)));
    const bool two = better || contract;  // scipy evaluates a second point
    if (two && left < 2) {  // ... but the budget ends after the reflection: the iteration is abandoned, nothing changes
        *fcalls += 1;
        return -2;
    }
    *fcalls += two ? 2 : 1;
    if (which >= 0) {
        const double fnew = which == 0 ? fxe : (which == 1 ? fxr : (which == 2 ? fxc : fxcc));
        int k = 0;
        for (int j = 0; j < N; ++j) k += fsim[j] <= fnew;  // the stable sort puts the new vertex behind its equals
        *rank = k;
        for (int i = 0; i < N; ++i) sim[N * N + i] = pts[which * N + i];
        fsim[N] = fnew;
        nm_sort(N, sim, fsim);
        *iters += 1;
        return which;
    }
    // shrink towards the best vertex: sim[0] + sigma (sim[j] - sim[0]); scipy moves and evaluates the vertices one at a
    // time, so with `rest` evaluations left vertices 1..rest + 1 move (the last of them is not evaluated any more)
    const long long rest = c.maxfev - *fcalls;
    const int nmove = rest < N ? (int)rest + 1 : N;
    for (int j = 1; j <= nmove; ++j)
        for (int k = 0; k < N; ++k) sim[j * N + k] = nm_add(sim[k], nm_mul(0.5, nm_sub(sim[j * N + k], sim[k])));
    if (rest >= 1) *phase = NM_SHRINK;  // else: nothing left to evaluate; the next round's budget test ends the fit
    return -1;
}

// The round's values fv_in[] (in the order of the submitted points; objective values, or llh = -objective with `negate`)
// applied to the simplex.
MISTI_HD inline void nm_apply(const NmConfig& c, double* sim, double* fsim, long long* iters, long long* fcalls, int* status,
                              int* phase, const double* pts, const double* fv_in, bool negate) {
    const int N = c.N;
    const long long left = c.maxfev - *fcalls;
    if (*phase == NM_INIT) {
        const int n = left < N + 1 ? (int)left : N + 1;
        for (int j = 0; j <= N; ++j) fsim[j] = j < n ? nm_clean(fv_in[j], negate) : kInf;  // past the budget: scipy's initial +inf
        *fcalls = n;
        *iters = 1;
        nm_sort(N, sim, fsim);
        *phase = NM_STEP;
        return;
    }
    if (*phase == NM_SHRINK) {
        const int n = left < N ? (int)left : N;
        for (int j = 1; j <= n; ++j) fsim[j] = nm_clean(fv_in[j - 1], negate);
        *fcalls += n;
        nm_sort(N, sim, fsim);
        if (n == N) *iters += 1;  // an iteration cut short by the budget is not counted (:929-931)
        *phase = NM_STEP;
        return;
    }
    int rank = 0;
    const int o = nm_step(c, sim, fsim, iters, fcalls, phase, pts, fv_in, negate, &rank);
    if (!c.lookahead || o < 0) return;
    // second iteration: the termination tests in between, then the scenario that came true
    if (nm_retire(c, sim, fsim, *iters, *fcalls, status)) {
        *phase = NM_DONE;
        return;
    }
    const int q = nm_scenario(N, o, rank);
    nm_step(c, sim, fsim, iters, fcalls, phase, pts + (long)(4 + 4 * q) * N, fv_in + 4 + 4 * q, negate, &rank);
}

// ------------------------------------------------------------------------------------------------
// Basin-hopping walkers (scipy/optimize/_basinhopping.py as MigrationInference.Solve(globalOpt=True) calls it,
// MigrationInference.py:724: niter = 100, T = 0.5, stepsize = 0.5, local search = Nelder-Mead with scipy's defaults).
// A walker is a Nelder-Mead simplex plus the state below; when its local search ends, bh_advance takes the Metropolis
// decision, updates the best-so-far storage and starts the next local search from a displaced point -- one walker at a
// time, with no barrier between the hops of different walkers (misti_b200/optim.py:basinhopping_batch is the same logic
// in lock step on the host; both reproduce scipy.optimize.basinhopping(..., rng=seed) for a single walker).
// Random numbers: numpy's Generator(PCG64) continued from the state the host seeded (numpy.random.default_rng(seed)),
// so the device draws the very numbers scipy would: uniform(-stepsize, stepsize, N) per displacement, uniform() per
// Metropolis test.
// ------------------------------------------------------------------------------------------------
struct Pcg64 { unsigned long long s_hi, s_lo, inc_hi, inc_lo; };

MISTI_HD inline unsigned long long mul64hi(unsigned long long a, unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}

// numpy/random/src/pcg64: state <- state * 0x2360ED051FC65DA44385DF649FCCF645 + inc (mod 2^128), output XSL-RR of the NEW state
MISTI_HD inline unsigned long long pcg64_next(Pcg64& r) {
    const unsigned long long m_hi = 0x2360ED051FC65DA4ull, m_lo = 0x4385DF649FCCF645ull;
    unsigned long long lo = r.s_lo * m_lo;
    unsigned long long hi = mul64hi(r.s_lo, m_lo) + r.s_hi * m_lo + r.s_lo * m_hi;
    const unsigned long long lo2 = lo + r.inc_lo;
    hi += r.inc_hi + (lo2 < lo ? 1ull : 0ull);
    r.s_lo = lo2; r.s_hi = hi;
    const unsigned long long x = hi ^ lo2;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
MISTI_HD inline double pcg64_double(Pcg64& r) { return (double)(pcg64_next(r) >> 11) * (1.0 / 9007199254740992.0); }
// Generator.uniform(low, high): low + (high - low) * next_double, every operation rounded on its own
MISTI_HD inline double pcg64_uniform(Pcg64& r, double low, double high) {
    return nm_add(low, nm_mul(nm_sub(high, low), pcg64_double(r)));
}

struct BhConfig {
    int niter;         // hops after the first local search; < 0 = plain Nelder-Mead fits (no walker state)
    int interval;      // the step size adapts every `interval` hops
    double beta;       // 1 / T (infinity for T = 0)
    double target, factor, stepsize0;
};

// one walker's state, as pointers into the structure-of-arrays block of the fit
struct BhWalker {
    double *x, *best_x;                 // [N]
    double *energy, *best_f, *step;     // scalars
    int *ok, *best_ok, *done;
    long long *nfev, *failures, *nstep, *naccept, *hop;
    Pcg64* rng;
};

// The walker's local search has ended (phase NM_DONE; sim / fsim sorted, status set).  Returns true when the walker goes on
// with another local search (sim[0] = the displaced start, phase NM_INIT), false when it has done its hops.
MISTI_HD inline bool bh_advance(const BhConfig& c, int N, const BhWalker& w, double* sim, double* fsim, long long* iters,
                                long long* fcalls, int* status, int* phase) {
    const double fn = fsim[0];
    const bool okn = *status == 0;
    if (*w.hop == 0) {  // the initial minimisation (_basinhopping.py: BasinHoppingRunner.__init__)
        for (int k = 0; k < N; ++k) { w.x[k] = sim[k]; w.best_x[k] = sim[k]; }
        *w.energy = fn; *w.ok = okn;
        *w.best_f = fn; *w.best_ok = okn;
        *w.nfev = *fcalls;
        *w.failures = okn ? 0 : 1;
    } else {  // Metropolis.accept_reject (the random number is drawn whatever the outcome), Storage.update
        *w.nfev += *fcalls;
        *w.failures += okn ? 0 : 1;
        const double prod = nm_mul(-nm_sub(fn, *w.energy), c.beta);
        const double wgt = exp(prod < 0.0 ? prod : 0.0);  // Python's min(0, prod): 0 when prod is NaN (inf - inf)
        const double u = pcg64_uniform(*w.rng, 0.0, 1.0);
        if (wgt >= u && (okn || !*w.ok)) {
            *w.naccept += 1;
            *w.energy = fn; *w.ok = okn;
            for (int k = 0; k < N; ++k) w.x[k] = sim[k];
            if (okn && (fn < *w.best_f || !*w.best_ok)) {
                *w.best_f = fn; *w.best_ok = okn;
                for (int k = 0; k < N; ++k) w.best_x[k] = sim[k];
            }
        }
    }
    *w.hop += 1;
    if (*w.hop > c.niter) { *w.done = 1; return false; }
    // AdaptiveStepsize.take_step: count, adapt every `interval` steps, displace
    *w.nstep += 1;
    if (*w.nstep % c.interval == 0) {
        const double rate = nm_div((double)*w.naccept, (double)*w.nstep);
        *w.step = rate > c.target ? nm_div(*w.step, c.factor) : nm_mul(*w.step, c.factor);
    }
    for (int k = 0; k < N; ++k) sim[k] = nm_add(w.x[k], pcg64_uniform(*w.rng, -*w.step, *w.step));
    for (int j = 0; j <= N; ++j) fsim[j] = 0.0;
    *iters = 0; *fcalls = 0; *status = -1; *phase = NM_INIT;
    return true;
}

}  // namespace misti
