/* misti_b200.h -- C ABI of libmisti_b200.so: batched MiSTI model evaluation on one B200 (sm_100a).
 *
 * The reference (Genomics-HSE/MiSTI) is pure Python and has no FFI layer; the boundary a replacement
 * has to honour is the Python class API.  This header declares the C entry
 * points the Python host in misti_b200/ binds with ctypes and which a maintainer of the reference
 * would bind in the same way (INTEGRATION.md shows the stub).  Each entry point cites the reference
 * interface it replaces.  Plain pointers and sizes only; no torch / numpy types.
 *
 * There is NO CPU path behind these functions: every evaluation runs in the CUDA kernels of
 * misti_b200/csrc/misti_kernels.cu, and misti_ctx_create fails when no CUDA device is usable.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MISTI_E_* code on API misuse or CUDA failure
 *     (text via misti_last_error); numerical outcomes are reported PER ITEM in status[] with the
 *     reference's conventions (llh = -inf where MigrationInference.JAFSLikelihood returns -inf).
 *   - the caller owns every buffer it passes; inputs are copied at call time.
 *   - a context is bound to one device and one stream; calls on one context must be serialised by
 *     the caller (the reference is single-threaded: MiSTI.py:23-25).  One context per process per GPU.
 */
#ifndef MISTI_B200_H
#define MISTI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MISTI_ABI_VERSION 2

/* error codes (function return values) */
#define MISTI_E_ARG (-1)     /* invalid argument / unknown id / capacity exceeded */
#define MISTI_E_CUDA (-2)    /* CUDA runtime failure (see misti_last_error)       */
#define MISTI_E_NODEV (-3)   /* no usable CUDA device: there is no CPU fallback   */

/* evaluation flags = the MigrationInference keyword arguments (MigrationInference.py:41-200) */
#define MISTI_FLAG_CORRECT 1u  /* not trueEPS: run the coalescence-rate correction (CorrectLambdas) */
#define MISTI_FLAG_CPFIT 2u    /* cpfit=True: fit non-coalescence probabilities                     */
#define MISTI_FLAG_SMOOTH 4u   /* smooth=True: SmoothConst over runs of equal PSMC rates            */
#define MISTI_FLAG_UNFOLDED 8u /* unfolded=True: 7-bin likelihood, else 4 folded bins               */
#define MISTI_FLAG_DEVICE_PTRS 256u /* params/model_ids/lc_inject/llh/jafs/status are DEVICE-ACCESSIBLE pointers; the call is
                                       asynchronous on the context's stream (no host synchronisation).  Pinned host
                                       memory qualifies (unified addressing): the kernels then read the parameters and
                                       write likelihoods, spectra and status across PCIe while they compute, and no
                                       copy trails the evaluation (misti_b200.parallel.ShardedEvaluator)              */

/* per-item status codes */
#define MISTI_OK 0
#define MISTI_NEGATIVE_PARAM 1     /* "Hit negative value of migration rate" (MigrationInference.py:569-572)  */
#define MISTI_CORRECTION_FAILED 2  /* "Lambda correction failed" (MigrationInference.py:575-578)              */
#define MISTI_NONFINITE 3          /* the reference would have raised or produced NaN                         */
#define MISTI_INFINITE_COAL_TIME 4 /* last interval before the split without migration (:475-476, ref. exits)  */
#define MISTI_SKIPPED 6            /* item whose model id is not a registered model (-1: slot left empty by the on-device
                                      optimiser; device-pointer calls are not validated on the host): llh = NaN      */
#define MISTI_STIFF 5              /* intervals WITH migration and (largest exit rate)*length > 256 are only seen after a
                                      run-away correction; they take a dense scaling-and-squaring step instead of
                                      the sweep (transparently: the item still ends with status 0).  The code is
                                      the transient state of such an item between the two kernels                */

#define MISTI_MAX_BANDS 8
#define MISTI_MAX_PULSES 8
#define MISTI_MAX_PARAMS 16

typedef struct misti_ctx misti_ctx;

/* One model layout = what MigrationInference.__init__ + SetModel hold (MigrationInference.py:41-200,
 * 229-289): the split interval, the sampling date of genome 2, migration bands [start, end) per
 * source deme and pulses, each either fixed (opt = -1) or bound to optimiser parameter `opt`
 * (MapParameters, :291-298: bands first, then pulses, in the order given).  pop is 0-based. */
typedef struct misti_model_desc {
    int32_t grid_id;     /* time grid registered with misti_add_grid                    */
    int32_t split_t;     /* first one-population interval (0 .. numT)                   */
    int32_t sample_date; /* interval index at which genome 2 was sampled (0 = present)  */
    int32_t n_bands;
    int32_t n_pulses;
    int32_t n_params;    /* number of optimiser parameters this model consumes          */
    int32_t band_pop[MISTI_MAX_BANDS], band_start[MISTI_MAX_BANDS], band_end[MISTI_MAX_BANDS], band_opt[MISTI_MAX_BANDS];
    int32_t pulse_pop[MISTI_MAX_PULSES], pulse_time[MISTI_MAX_PULSES], pulse_opt[MISTI_MAX_PULSES];
    double band_val[MISTI_MAX_BANDS];
    double pulse_val[MISTI_MAX_PULSES];
} misti_model_desc;

/* Optional outputs / inputs of misti_eval_batch; any pointer may be NULL.  Host or device pointers
 * according to MISTI_FLAG_DEVICE_PTRS.  numT_max = largest numT over the registered grids. */
typedef struct misti_eval_io {
    const double* lc_inject; /* [B][numT_max][2] corrected rates to use instead of running the correction   */
    double* jafs;            /* [B][7]  normalised expected JSFS (MigrationInference.JAFS)                  */
    double* jafs_raw;        /* [B][7]  JAFSpectrum() before normalisation                                  */
    double* lc_out;          /* [B][numT_max][2] corrected rates (MigrationInference.lc)                    */
    double* pr_out;          /* [B][numT_max+1][3][2] 3-state trajectories (MigrationInference.Pr)          */
    int32_t* status;         /* [B]                                                                        */
    int32_t* nfev;           /* [B] residual evaluations spent in the correction (least_squares nfev sum)  */
    int32_t* terms;          /* [B] sparse mat-vecs spent in the JSFS stage (a closed-form zero-migration run = 1) */
    const int32_t* row_ids;  /* [B] score item b against data row row_ids[b] ONLY; llh is then [B] instead of
                                [B][R] (one optimiser simplex per (bootstrap row, split time) pair)         */
    int32_t* solve_trace;    /* [B][numT_max][2] per interval: evaluations (`nfev`) and termination `status` of its
                                scipy.optimize.least_squares solve (CorrectLambda.py:85, 260, 303, 305), (0, -9) where
                                the interval has a closed form -- the iterate-level record behind `nfev`    */
    double* row_best_llh;    /* [R] max over the B items of llh[b][r], reduced ON THE DEVICE, and                   */
    int32_t* row_best_item;  /* [R] the item that attains it (the first one): what the reference's bootstrap notebook
                                computes from one result line per (row, split time) (test.bs/bs_conf_int.ipynb).  With
                                these set `llh` may be NULL: the B x R likelihoods then never leave the device.  Host
                                pointers, no row_ids                                                          */
} misti_eval_io;

int misti_abi_version(void);

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t to launch on (e.g. the
 * caller's torch stream) or NULL for a private stream.  Replaces: construction of the model
 * objects in MiSTI.py:213. */
int misti_ctx_create(int device, void* stream, misti_ctx** out);
void misti_ctx_destroy(misti_ctx* ctx);
const char* misti_last_error(const misti_ctx* ctx);
int misti_ctx_set_stream(misti_ctx* ctx, void* stream);
int misti_ctx_synchronize(misti_ctx* ctx);
/* Optional: size the context's scratch buffers for batches of up to B items with P parameters and rows_per_item
 * likelihoods each (1 with misti_eval_io.row_ids, else the number of data rows), for the grids and models registered so
 * far, so that a run whose batches grow does not re-allocate on the way. */
int misti_ctx_reserve(misti_ctx* ctx, int32_t B, int32_t P, int32_t rows_per_item);

/* Register a merged PSMC time grid: times[numT-1] interval lengths and lh[numT][2] apparent
 * coalescence rates of the two genomes (InputData.times / .lambdas from migrationIO.ReadPSMC,
 * migrationIO.py:224-295, after the constructor's fractional-split surgery, MigrationInference.py:89-99). */
int misti_add_grid(misti_ctx* ctx, int32_t numT, const double* times, const double* lh, int32_t* grid_id);

/* Register a model layout (MigrationInference.SetModel, MigrationInference.py:229-289). */
int misti_add_model(misti_ctx* ctx, const misti_model_desc* desc, int32_t* model_id);
/* Drop all grids and models (data rows are kept). */
int misti_clear_models(misti_ctx* ctx);

/* Observed spectra: sfs[R][8] = [total sites, 7 counts] rows as read by MiSTI.py:172-178 (row 0 =
 * data, further rows = bootstrap replicates).  llh_const[R] = lnGamma(n+1) - sum lnGamma(k_i+1)
 * over the 7 (unfolded) or 4 folded bins (MigrationInference.SetJAFS, :202-227); NULL = computed
 * here with lgamma().  `unfolded` selects the binning. */
int misti_set_data(misti_ctx* ctx, int32_t R, const double* sfs, const double* llh_const, int32_t unfolded);

/* Evaluate B items.  Item b uses model model_ids[b] (or `model_default` when model_ids is NULL)
 * and the optimiser vector params[b*P .. b*P+P) (P >= that model's n_params, P <= MISTI_MAX_PARAMS).
 * llh[b*R + r] = composite log-likelihood of item b against data row r (MigrationInference
 * .JAFSLikelihood, MigrationInference.py:566-614); -inf where the reference returns -inf.
 * Replaces: one MigrationInference.JAFSLikelihood call per item per data row. */
int misti_eval_batch(misti_ctx* ctx, int32_t B, int32_t P, const double* params, const int32_t* model_ids,
                     int32_t model_default, uint32_t flags, double mixture_th, double* llh, const misti_eval_io* io);

/* Fit S independent (model, data row) pairs by Nelder-Mead ON THE DEVICE: every simplex takes the decisions of
 * scipy.optimize.minimize(method='Nelder-Mead') as MigrationInference.Solve calls it (MigrationInference.py:718-729;
 * scipy 1.18.1 _optimize.py:_minimize_neldermead), the objective is -llh of the pair's model against its data row, and
 * a fit is a stream of launches (propose, evaluate, apply) without a host round trip per step.  Host pointers.
 * x0[S][N] start vectors (N >= every model's n_params, N >= 1), model_ids[S], row_ids[S] (NULL = row 0);
 * maxiter / maxfev < 0 = none.  Out: x[S][N] best vertex, fun[S] = -llh there, nit[S], nfev[S] (scipy's counts),
 * status[S] (0 converged, 1 maxfev, 2 maxiter), info[3] = rounds of launches, points evaluated, 1 if the rounds were
 * replayed as a CUDA graph (nullable).
 * Replaces: one MigrationInference.Solve (one MiSTI.py process in the reference's bootstrap loops) per pair. */
int misti_nelder_mead(misti_ctx* ctx, int32_t S, int32_t N, const double* x0, const int32_t* model_ids, const int32_t* row_ids,
                      uint32_t flags, double mixture_th, double xatol, double fatol, int64_t maxiter, int64_t maxfev,
                      double* x, double* fun, int64_t* nit, int64_t* nfev, int32_t* status, int64_t* info);

/* The general form of the on-device optimiser: S independent fits, each a Nelder-Mead simplex that takes scipy's decisions,
 * optionally wrapped in a basin-hopping walker (scipy.optimize.basinhopping as MigrationInference.Solve(globalOpt=True) calls
 * it, MigrationInference.py:724: niter = 100, T = 0.5, stepsize = 0.5, local search = Nelder-Mead with scipy's defaults
 * xatol = fatol = 1e-4, maxiter = maxfev = 200 N).  Nothing waits for the slowest fit: the points of a round are packed
 * behind a device-side counter, the evaluation kernels read the count from the device, and a walker whose local search has
 * ended takes its Metropolis decision and starts its next local search in the following round, whatever the other walkers
 * are doing.  A single walker with generator state s reproduces scipy.optimize.basinhopping(..., rng=s).  Host pointers. */
typedef struct misti_fit_opts {
    double xatol, fatol;       /* Nelder-Mead termination (of every local search)                                   */
    int64_t maxiter, maxfev;   /* per local search; < 0 = none                                                       */
    int32_t niter;             /* basin-hopping hops after the initial minimisation; < 0 = plain Nelder-Mead fits    */
    int32_t interval;          /* the step size adapts every `interval` hops (scipy: 50)                             */
    double T, stepsize, target_accept_rate, stepwise_factor; /* scipy: 1.0 (reference: 0.5), 0.5, 0.5, 0.9          */
    const uint64_t* rng_state; /* [S][4] numpy PCG64 state per walker: state >> 64, state & (2^64 - 1), inc >> 64,
                                  inc & (2^64 - 1) of numpy.random.default_rng(seed).bit_generator.state             */
} misti_fit_opts;
typedef struct misti_fit_result {
    double* x;        /* [S][N] plain fits: best vertex; walkers: lowest minimum seen (scipy's Storage)              */
    double* fun;      /* [S]    -llh there                                                                           */
    int64_t* nit;     /* [S]    plain fits: iterations; walkers: hops taken                                          */
    int64_t* nfev;    /* [S]    scipy's evaluation count (walkers: summed over the local searches)                   */
    int32_t* status;  /* [S]    plain fits: 0 converged, 1 maxfev, 2 maxiter; walkers: 0 = the minimum is a converged one */
    int64_t* accepted;/* [S]    walkers: accepted hops (nullable)                                                    */
    int64_t* failures;/* [S]    walkers: local searches that did not converge (nullable)                             */
    int64_t rounds, points; /* out: rounds of launches that evaluated something, points evaluated                    */
    int32_t graph;    /* out: 1 if the rounds were replayed as a CUDA graph                                          */
} misti_fit_result;
int misti_fit(misti_ctx* ctx, int32_t S, int32_t N, const double* x0, const int32_t* model_ids, const int32_t* row_ids,
              uint32_t flags, double mixture_th, const misti_fit_opts* opts, misti_fit_result* res);

/* Score B given spectra (7 non-negative weights each, normalised on the device) against every data
 * row: llh[b*R + r].  Host pointers.  Replaces the likelihood tail used on its own, e.g.
 * MigrationInference.MaximumLLHFunction (MigrationInference.py:696-711) with the data spectrum. */
int misti_score_spectra(misti_ctx* ctx, int32_t B, const double* spectra, double* llh);

/* The forward map true rates -> PSMC-apparent rates, MigrationInference.CoalescentRates (MigrationInference.py:542-564;
 * CorrectLambda.CoalRates, CorrectLambda.py:112-122), used by TestModel.py:120 and to write model-consistent PSMC input:
 * the registered grid's rates are taken as the TRUE rates of model `model_id`.  mu0, mu1: the reference propagates every
 * interval with the migration rates its CorrectLambda helper was left with by the preceding likelihood call (those of
 * interval split_t - 1); the caller passes them.  Out (host pointers): lh_out[numT][2] and, nullable,
 * pr_out[(min(split_t, numT) + 1)][3][2], the trajectory of the two 3-state chains (.Pr). */
int misti_coalescent_rates(misti_ctx* ctx, int32_t model_id, int32_t P, const double* params, double mu0, double mu1,
                           double* lh_out, double* pr_out);

/* Device time in milliseconds of the two kernels of the last misti_eval_batch on this context
 * (CUDA events on the context's stream): out[0] = correction kernel, out[1] = JSFS+likelihood kernel.
 * Synchronises with the stream. */
int misti_last_kernel_ms(misti_ctx* ctx, float* out2);
/* Number of kernel launches issued by this context so far. */
int64_t misti_launch_count(const misti_ctx* ctx);

/* Structure tables of the lineage chains, produced ON THE DEVICE from the same tables the kernels use
 * (for the TwoPopulations / OnePopulation mirror classes).  generator: 44x44 row-major
 * (TwoPopulations.SetMatrix before the stationary states are deleted, TwoPopulations.py:231-238);
 * which = 1: one-population 8x8 generator for rate l1 (OnePopulation.SetMatrix, OnePopulation.py:153-158). */
int misti_generator(misti_ctx* ctx, int32_t which, double l1, double l2, double m1, double m2, double* out);
/* PulseMigration (TwoPopulations.py:361-377) and AncientSampleP0 (:246-262) applied to P0[44]. */
int misti_pulse(misti_ctx* ctx, const double* P0, double rate, int32_t src_pop, double* P1);
int misti_ancient_reset(misti_ctx* ctx, const double* P0, double* P1);
/* StateToJAF for all states: out[44][7] (which = 0) or out[8][7] (which = 1). */
int misti_state_to_jaf(misti_ctx* ctx, int32_t which, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* MISTI_B200_H */
