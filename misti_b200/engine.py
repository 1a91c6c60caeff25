"""Engine: one CUDA context of libmisti_b200.so on one GPU, as a small Python object.

Thin host-side plumbing over the C ABI (include/misti_b200.h): registers merged PSMC grids, model
layouts (split time, migration bands, pulses) and observed spectra, and evaluates batches of
optimiser vectors.  All arithmetic of the evaluation happens in the CUDA kernels; nothing here
computes a likelihood on the host.
"""
import ctypes

import numpy as np

from . import _lib

STATUS_TEXT = {
    _lib.OK: "ok",
    _lib.NEGATIVE_PARAM: "negative parameter",
    _lib.CORRECTION_FAILED: "lambda correction failed",
    _lib.NONFINITE: "non-finite result",
    _lib.INFINITE_COAL_TIME: "infinite coalescent time, no migration",
    _lib.STIFF: "interval too stiff (run-away corrected rate)",
}


def _as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def llh_constants(rows, unfolded):
    """lnGamma(n+1) - sum lnGamma(k_i+1) per data row, with scipy.special.gammaln exactly as
    MigrationInference.SetJAFS does (MigrationInference.py:216-227): the same terms subtracted in the same
    order, vectorised over the rows (1001 bootstrap rows cost 0.2 s one scalar call at a time)."""
    from scipy.special import gammaln
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    d = rows[:, 1:]
    n = np.zeros(len(rows))
    for i in range(7):  # Python's sum(): left to right, starting from 0
        n = n + d[:, i]
    c = 0 + gammaln(n + 1)
    if unfolded:
        for i in range(7):
            c = c - gammaln(d[:, i] + 1)
    else:
        c = c - gammaln(d[:, 0] + d[:, 6] + 1)
        c = c - gammaln(d[:, 1] + d[:, 5] + 1)
        c = c - gammaln(d[:, 2] + d[:, 4] + 1)
        c = c - gammaln(d[:, 3] + 1)
    return np.asarray(c, dtype=np.float64).reshape(-1)


class Engine:
    """One device context.  Not thread-safe; use one Engine per process per GPU."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.misti_ctx_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if rc == _lib.E_NODEV:
            raise _lib.MistiLibraryError("no usable CUDA device %d: misti_b200 has no CPU fallback" % device)
        if rc != 0:
            raise _lib.MistiLibraryError("misti_ctx_create failed (%d)" % rc)
        self._h = h
        self.device = int(device)
        self.numT_max = 0
        self.grids = []   # numT per grid
        self.models = []  # dict per model
        self.R = 0
        self.unfolded = True

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.misti_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.misti_last_error(self._h)
            raise _lib.MistiLibraryError("libmisti_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def set_stream(self, stream):
        self._check(self._lib.misti_ctx_set_stream(self._h, ctypes.c_void_p(stream) if stream else None))

    def synchronize(self):
        self._check(self._lib.misti_ctx_synchronize(self._h))

    def reserve(self, B, P=1, rows_per_item=1):
        """Size the scratch buffers for batches of up to B items (optional; avoids re-allocation while batches grow).
        rows_per_item: likelihoods per item -- 1 for calls with row_ids, else the number of data rows."""
        self._check(self._lib.misti_ctx_reserve(self._h, int(B), int(P), int(rows_per_item)))

    # -- registration ---------------------------------------------------------------------------
    def add_grid(self, times, lambdas):
        lh = _as_f64(lambdas).reshape(-1, 2)
        numT = lh.shape[0]
        t = _as_f64(times).reshape(-1)
        if t.shape[0] != numT - 1:
            raise ValueError("Unexpected number of time intervals")
        gid = ctypes.c_int32()
        self._check(self._lib.misti_add_grid(self._h, numT, t.ctypes.data_as(_lib.c_double_p),
                                             lh.ctypes.data_as(_lib.c_double_p), ctypes.byref(gid)))
        self.grids.append(numT)
        self.numT_max = max(self.numT_max, numT)
        return gid.value

    def add_model(self, grid_id, splitT, sampleDate=0, bands=(), pulses=()):
        """bands: (pop0, start, end, value, opt) with opt = optimiser index or -1; pulses: (pop0, time, value, opt)."""
        d = _lib.ModelDesc()
        d.grid_id, d.split_t, d.sample_date = int(grid_id), int(splitT), int(sampleDate)
        if len(bands) > _lib.MAX_BANDS or len(pulses) > _lib.MAX_PULSES:
            raise ValueError("too many migration bands / pulses (max %d / %d)" % (_lib.MAX_BANDS, _lib.MAX_PULSES))
        d.n_bands, d.n_pulses = len(bands), len(pulses)
        n_params = 0
        for i, (pop, a, b, val, opt) in enumerate(bands):
            d.band_pop[i], d.band_start[i], d.band_end[i], d.band_val[i], d.band_opt[i] = int(pop), int(a), int(b), float(val), int(opt)
            n_params = max(n_params, int(opt) + 1)
        for i, (pop, t, val, opt) in enumerate(pulses):
            d.pulse_pop[i], d.pulse_time[i], d.pulse_val[i], d.pulse_opt[i] = int(pop), int(t), float(val), int(opt)
            n_params = max(n_params, int(opt) + 1)
        d.n_params = n_params
        mid = ctypes.c_int32()
        self._check(self._lib.misti_add_model(self._h, ctypes.byref(d), ctypes.byref(mid)))
        self.models.append(dict(grid=int(grid_id), numT=self.grids[grid_id], splitT=int(splitT), n_params=n_params))
        return mid.value

    def clear_models(self):
        self._check(self._lib.misti_clear_models(self._h))
        self.grids, self.models, self.numT_max = [], [], 0

    def set_data(self, rows, unfolded, llh_const=None):
        """rows[R][8] = [total sites, 7 counts]; row 0 is the data, further rows bootstrap replicates."""
        rows = _as_f64(rows).reshape(-1, 8)
        if llh_const is None:
            llh_const = llh_constants(rows, unfolded)
        c = _as_f64(llh_const).reshape(-1)
        if c.shape[0] != rows.shape[0]:
            raise ValueError("llh_const must have one entry per data row")
        self._check(self._lib.misti_set_data(self._h, rows.shape[0], rows.ctypes.data_as(_lib.c_double_p),
                                             c.ctypes.data_as(_lib.c_double_p), 1 if unfolded else 0))
        self.R = rows.shape[0]
        self.unfolded = bool(unfolded)

    # -- evaluation -----------------------------------------------------------------------------
    def evaluate(self, params, model=0, model_ids=None, flags=_lib.FLAG_CORRECT, mixtureTH=0.0, lc_inject=None,
                 want=("jafs", "status"), buffers=None, row_ids=None, row_best=False):
        """Host-buffer evaluation.  params: [B, P] (or [B] / [] for P = 0).  Returns a dict with
        'llh' [B, R] and the arrays named in `want` (jafs, jafs_raw, lc, pr, status, nfev, terms, solve_trace).
        `buffers` may hold preallocated C-contiguous numpy arrays (e.g. views of pinned memory) to write into.
        With `row_ids` [B], item b is scored against data row row_ids[b] only and 'llh' is [B, 1].
        row_best: reduce over the items ON THE DEVICE -- 'row_best_llh' [R] = max_b llh[b, r] and 'row_best_item' [R] = the item
        that attains it; with row_best="only" the [B, R] likelihoods never leave the device ('llh' is then absent)."""
        params = _as_f64(params)
        if params.ndim == 1:
            params = params.reshape(-1, 1) if params.size else params.reshape(1, 0)
        B, P = params.shape
        mids = None
        if model_ids is not None:
            mids = np.ascontiguousarray(model_ids, dtype=np.int32).reshape(-1)
            if mids.shape[0] != B:
                raise ValueError("model_ids must have one entry per item")
        flags = int(flags) & ~_lib.FLAG_DEVICE_PTRS
        flags = (flags | _lib.FLAG_UNFOLDED) if self.unfolded else (flags & ~_lib.FLAG_UNFOLDED)
        buffers = buffers or {}

        def _buf(name, shp, dt):
            a = buffers.get(name)
            if a is None:
                return np.zeros(shp, dtype=dt)
            if a.dtype != dt or a.size != int(np.prod(shp)) or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("buffer %r has the wrong dtype/size/layout" % name)
            return a.reshape(shp)
        io = _lib.EvalIO()
        rows = None
        if row_ids is not None:
            rows = np.ascontiguousarray(row_ids, dtype=np.int32).reshape(-1)
            if rows.shape[0] != B or (B and (rows.min() < 0 or rows.max() >= self.R)):
                raise ValueError("row_ids must hold one valid data row index per item")
            io.row_ids = _ptr(rows)
        out = {} if row_best == "only" else {"llh": _buf("llh", (B, 1 if rows is not None else self.R), np.float64)}
        if row_best:
            out["row_best_llh"], out["row_best_item"] = np.zeros(self.R), np.zeros(self.R, dtype=np.int32)
            io.row_best_llh, io.row_best_item = _ptr(out["row_best_llh"]), _ptr(out["row_best_item"])
        nT = self.numT_max
        if lc_inject is not None:
            inj = _as_f64(lc_inject).reshape(B, nT, 2)
            io.lc_inject = _ptr(inj)
        shapes = {"jafs": ((B, 7), np.float64), "jafs_raw": ((B, 7), np.float64), "lc": ((B, nT, 2), np.float64),
                  "pr": ((B, nT + 1, 3, 2), np.float64), "status": ((B,), np.int32), "nfev": ((B,), np.int32),
                  "terms": ((B,), np.int32), "solve_trace": ((B, nT, 2), np.int32)}
        field = {"jafs": "jafs", "jafs_raw": "jafs_raw", "lc": "lc_out", "pr": "pr_out", "status": "status", "nfev": "nfev",
                 "terms": "terms", "solve_trace": "solve_trace"}
        for name in want:
            shp, dt = shapes[name]
            out[name] = _buf(name, shp, dt)
            setattr(io, field[name], _ptr(out[name]))
        self._check(self._lib.misti_eval_batch(self._h, B, P, _ptr(params) if P else None, _ptr(mids), int(model), flags,
                                               float(mixtureTH), _ptr(out.get("llh")), ctypes.byref(io)))
        return out

    def evaluate_device(self, B, P, params_ptr, llh_ptr, model=0, model_ids_ptr=None, flags=_lib.FLAG_CORRECT, mixtureTH=0.0,
                        jafs_ptr=None, status_ptr=None, terms_ptr=None, nfev_ptr=None):
        """Asynchronous evaluation on device-resident buffers (raw device addresses, e.g. torch
        tensor .data_ptr()); work is queued on the context's stream and NOT synchronised."""
        flags = int(flags) | _lib.FLAG_DEVICE_PTRS
        flags = (flags | _lib.FLAG_UNFOLDED) if self.unfolded else (flags & ~_lib.FLAG_UNFOLDED)
        io = _lib.EvalIO()
        io.jafs, io.status, io.terms, io.nfev = jafs_ptr, status_ptr, terms_ptr, nfev_ptr
        self._check(self._lib.misti_eval_batch(self._h, int(B), int(P), ctypes.c_void_p(params_ptr) if params_ptr else None,
                                               ctypes.c_void_p(model_ids_ptr) if model_ids_ptr else None, int(model), flags,
                                               float(mixtureTH), ctypes.c_void_p(llh_ptr), ctypes.byref(io)))

    def nelder_mead(self, x0, model_ids, row_ids=None, flags=_lib.FLAG_CORRECT, mixtureTH=0.0, xatol=1e-4, fatol=1e-4,
                    maxiter=None, maxfev=None):
        """Nelder-Mead fits of S (model, data row) pairs on the device (misti_nelder_mead): the objective is -llh, every
        simplex takes scipy's decisions, and no step returns to the host.  x0: [S, N].  maxiter / maxfev as in
        scipy.optimize.minimize (both None: N * 200 each).  Returns the dict of misti_b200.optim.nelder_mead_batch
        (x, fun, nit, nfev, status, success, evaluations, launches = rounds of device launches)."""
        x0 = _as_f64(x0)
        if x0.ndim == 1:
            x0 = x0.reshape(1, -1)
        S, N = x0.shape
        mids = np.ascontiguousarray(model_ids, dtype=np.int32).reshape(-1)
        rows = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int32).reshape(-1)
        if mids.shape[0] != S or (rows is not None and rows.shape[0] != S):
            raise ValueError("model_ids / row_ids must have one entry per start vector")
        if maxiter is None and maxfev is None:  # scipy's defaults (_optimize.py:751-768)
            maxiter, maxfev = N * 200, N * 200
        elif maxiter is None:
            maxiter = N * 200 if maxfev == np.inf else np.inf
        elif maxfev is None:
            maxfev = N * 200 if maxiter == np.inf else np.inf
        lim = [(-1 if v == np.inf else int(v)) for v in (maxiter, maxfev)]
        flags = int(flags) & ~_lib.FLAG_DEVICE_PTRS
        flags = (flags | _lib.FLAG_UNFOLDED) if self.unfolded else (flags & ~_lib.FLAG_UNFOLDED)
        x, fun = np.empty((S, N)), np.empty(S)
        nit, nfev, info = np.zeros(S, dtype=np.int64), np.zeros(S, dtype=np.int64), np.zeros(3, dtype=np.int64)
        status = np.zeros(S, dtype=np.int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        self._check(self._lib.misti_nelder_mead(
            self._h, S, N, x0.ctypes.data_as(_lib.c_double_p), mids.ctypes.data_as(_lib.c_int32_p),
            rows.ctypes.data_as(_lib.c_int32_p) if rows is not None else None, flags, float(mixtureTH), float(xatol), float(fatol),
            lim[0], lim[1], x.ctypes.data_as(_lib.c_double_p), fun.ctypes.data_as(_lib.c_double_p), nit.ctypes.data_as(i64p),
            nfev.ctypes.data_as(i64p), status.ctypes.data_as(_lib.c_int32_p), info.ctypes.data_as(i64p)))
        return {"x": x, "fun": fun, "nit": nit, "nfev": nfev, "status": status.astype(np.int64), "success": status == 0,
                "evaluations": int(info[1]), "launches": int(info[0]), "graph": bool(info[2])}

    def basinhopping(self, x0, model_ids, row_ids=None, seeds=None, flags=_lib.FLAG_CORRECT, mixtureTH=0.0, niter=100, T=0.5,
                     stepsize=0.5, interval=50, target_accept_rate=0.5, stepwise_factor=0.9, xatol=1e-4, fatol=1e-4,
                     maxiter=None, maxfev=None):
        """W basin-hopping walkers on the device (misti_fit), as MigrationInference.Solve(globalOpt=True) runs ONE
        (MigrationInference.py:724: niter = 100, T = 0.5, stepsize = 0.5; local search = Nelder-Mead with scipy's defaults).
        No walker waits for another: a walker whose local search has ended takes its Metropolis decision and starts the next
        one in the following round of launches.  seeds: one seed (or numpy Generator) per walker -- walker w draws the numbers of
        numpy.random.default_rng(seeds[w]), so a single walker reproduces scipy.optimize.basinhopping(..., rng=seeds[w]).
        Returns the dict of misti_b200.optim.basinhopping_batch (x, fun, success, nfev, nit, accepted,
        minimization_failures, evaluations, launches)."""
        x0 = _as_f64(x0)
        if x0.ndim == 1:
            x0 = x0.reshape(1, -1)
        W, N = x0.shape
        mids = np.ascontiguousarray(model_ids, dtype=np.int32).reshape(-1)
        rows = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int32).reshape(-1)
        if mids.shape[0] != W or (rows is not None and rows.shape[0] != W):
            raise ValueError("model_ids / row_ids must have one entry per walker")
        seeds = list(range(W)) if seeds is None else list(seeds)
        if len(seeds) != W:
            raise ValueError("one seed per walker")
        state = np.empty((W, 4), dtype=np.uint64)
        mask = (1 << 64) - 1
        for w, sd in enumerate(seeds):  # numpy seeds the generator (SeedSequence -> PCG64); the device continues its stream
            g = sd if isinstance(sd, np.random.Generator) else np.random.default_rng(sd)
            st = g.bit_generator.state
            if st["bit_generator"] != "PCG64":
                raise ValueError("walker generators must be PCG64 (numpy.random.default_rng)")
            v, inc = int(st["state"]["state"]), int(st["state"]["inc"])
            state[w] = (v >> 64, v & mask, inc >> 64, inc & mask)
        if maxiter is None and maxfev is None:  # scipy's defaults (_optimize.py:751-768)
            maxiter, maxfev = N * 200, N * 200
        elif maxiter is None:
            maxiter = N * 200 if maxfev == np.inf else np.inf
        elif maxfev is None:
            maxfev = N * 200 if maxiter == np.inf else np.inf
        o = _lib.FitOpts()
        o.xatol, o.fatol = float(xatol), float(fatol)
        o.maxiter, o.maxfev = [(-1 if v == np.inf else int(v)) for v in (maxiter, maxfev)]
        o.niter, o.interval, o.T, o.stepsize = int(niter), int(interval), float(T), float(stepsize)
        o.target_accept_rate, o.stepwise_factor = float(target_accept_rate), float(stepwise_factor)
        o.rng_state = state.ctypes.data_as(ctypes.c_void_p)
        x, fun = np.empty((W, N)), np.empty(W)
        nit, nfev, acc, fail = (np.zeros(W, dtype=np.int64) for _ in range(4))
        status = np.zeros(W, dtype=np.int32)
        r = _lib.FitResult()
        r.x, r.fun, r.nit, r.nfev, r.status, r.accepted, r.failures = (_ptr(a) for a in (x, fun, nit, nfev, status, acc, fail))
        flags = int(flags) & ~_lib.FLAG_DEVICE_PTRS
        flags = (flags | _lib.FLAG_UNFOLDED) if self.unfolded else (flags & ~_lib.FLAG_UNFOLDED)
        self._check(self._lib.misti_fit(self._h, W, N, x0.ctypes.data_as(_lib.c_double_p), mids.ctypes.data_as(_lib.c_int32_p),
                                        rows.ctypes.data_as(_lib.c_int32_p) if rows is not None else None, flags, float(mixtureTH),
                                        ctypes.byref(o), ctypes.byref(r)))
        return {"x": x, "fun": fun, "success": status == 0, "nfev": nfev, "nit": nit, "accepted": acc,
                "minimization_failures": fail, "evaluations": int(r.points), "launches": int(r.rounds), "graph": bool(r.graph)}

    def coalescent_rates(self, model, params, mu):
        """Forward map of model `model` (misti_coalescent_rates): its grid's rates taken as the true rates -> the rates PSMC
        would see, lh [numT, 2], and the chain trajectory Pr [splitT + 1, 3, 2].  mu = (mu0, mu1), see the header."""
        m = self.models[int(model)]
        numT, n2 = m["numT"], min(m["splitT"], m["numT"])
        p = _as_f64(params).reshape(-1)
        lh, pr = np.zeros((numT, 2)), np.zeros((n2 + 1, 3, 2))
        self._check(self._lib.misti_coalescent_rates(self._h, int(model), int(p.shape[0]), p.ctypes.data_as(_lib.c_double_p) if p.size else None,
                                                     float(mu[0]), float(mu[1]), lh.ctypes.data_as(_lib.c_double_p),
                                                     pr.ctypes.data_as(_lib.c_double_p)))
        return lh, pr

    def score_spectra(self, spectra):
        """llh [B, R] of given spectra (7 weights each, normalised on the device) against every data row."""
        sp = _as_f64(spectra).reshape(-1, 7)
        out = np.empty((sp.shape[0], self.R))
        self._check(self._lib.misti_score_spectra(self._h, sp.shape[0], sp.ctypes.data_as(_lib.c_double_p),
                                                  out.ctypes.data_as(_lib.c_double_p)))
        return out

    def last_kernel_ms(self):
        """(correction kernel ms, JSFS+likelihood kernel ms) of the last evaluation; synchronises."""
        buf = (ctypes.c_float * 2)()
        self._check(self._lib.misti_last_kernel_ms(self._h, buf))
        return float(buf[0]), float(buf[1])

    def launch_count(self):
        return int(self._lib.misti_launch_count(self._h))

    # -- structure tables (device-produced) -----------------------------------------------------
    def generator(self, l1, l2=0.0, m1=0.0, m2=0.0, one_pop=False):
        n = 8 if one_pop else 44
        out = np.zeros((n, n))
        self._check(self._lib.misti_generator(self._h, 1 if one_pop else 0, float(l1), float(l2), float(m1), float(m2),
                                              out.ctypes.data_as(_lib.c_double_p)))
        return out

    def pulse(self, P0, rate, src_pop):
        P0 = _as_f64(P0).reshape(44)
        out = np.zeros(44)
        self._check(self._lib.misti_pulse(self._h, P0.ctypes.data_as(_lib.c_double_p), float(rate), int(src_pop),
                                          out.ctypes.data_as(_lib.c_double_p)))
        return out

    def ancient_reset(self, P0):
        P0 = _as_f64(P0).reshape(44)
        out = np.zeros(44)
        self._check(self._lib.misti_ancient_reset(self._h, P0.ctypes.data_as(_lib.c_double_p), out.ctypes.data_as(_lib.c_double_p)))
        return out

    def state_to_jaf(self, one_pop=False):
        n = 8 if one_pop else 44
        out = np.zeros((n, 7), dtype=np.int32)
        self._check(self._lib.misti_state_to_jaf(self._h, 1 if one_pop else 0, out.ctypes.data_as(_lib.c_int32_p)))
        return out


_default_engines = {}


def default_engine(device=0):
    """Process-wide Engine per device (created on first use)."""
    eng = _default_engines.get(device)
    if eng is None or eng._h is None:
        eng = Engine(device)
        _default_engines[device] = eng
    return eng
