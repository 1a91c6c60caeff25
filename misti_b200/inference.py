"""MigrationInference: drop-in for the reference class of the same name, evaluated on the GPU.

Mirrors the public surface of MigrationInference (reference MigrationInference.py:35-739):
constructor signature and keyword arguments, SetJAFS, SetModel, MapParameters, CorrectLambdas,
JAFSpectrum, JAFSLikelihood, ObjectiveFunction, MaximumLLHFunction, Solve, Report, and the
attributes other code reads (lh, lc, mi, pu, times, splitT, sampleDate, thrh, JAFS, dataJAFS,
Pr, llh, llh_const, snps).  Everything numerical (correction chain, expected JSFS, likelihood) runs
in libmisti_b200.so's CUDA kernels through misti_b200.engine.Engine; this module is host-side
bookkeeping only and raises when the CUDA library or a GPU is missing.

Additions over the reference (batched entry points):
    JAFSLikelihoodBatch(params[B, P]) -> llh[B] (or [B, R] with extra data rows)
    ExpectedJAFSBatch(params[B, P])   -> jafs[B, 7]
    SetJAFSBatch(rows[R, 8])          -> score every evaluation against R spectra at once
"""
import sys

import numpy as np

from . import _lib
from .engine import Engine


class MigrationInference:
    COUNT_LLH = 0
    CORRECTION_CALLED = 0
    CORRECTION_FAILED = 0

    def __init__(self, times, lambdas, dataJAFS, splitT, mi=[], pu=[], **kwargs):
        self.debug = bool(kwargs.get("debug", False))
        self.enableOutput = self.debug or bool(kwargs.get("enableOutput", False))
        if self.enableOutput:
            print("MigrationInference: output enabled.")
        self.cpfit = bool(kwargs.get("cpfit", False))
        self.correct = not bool(kwargs.get("trueEPS", False))
        self.smooth = bool(kwargs.get("smooth", False))
        self.LLHpsmc = "Tpsmc" in kwargs
        if self.LLHpsmc:
            self.Tpsmc = kwargs["Tpsmc"]
        self.unfolded = bool(kwargs.get("unfolded", False))
        self.thrh = [1.0, 1.0]
        if "thrh" in kwargs and len(kwargs["thrh"]) == 2:
            self.thrh = kwargs["thrh"]
        self.sampleDate = kwargs.get("sampleDate", 0)
        self.mixtureTH = float(kwargs.get("mixtureTH", 0.0))
        self.doPlot = bool(kwargs.get("doPlot", False))
        self._device = int(kwargs.get("device", 0))
        self._engine = kwargs.get("engine", None)
        self._owns_engine = self._engine is None
        self._registered = False

        if splitT < self.sampleDate:
            self.PrintError("__init__", "cannot initialise class with split time being more recent than sample date.")
        # fractional split time: the interval int(splitT) is cut in two IN THE CALLER'S LISTS, exactly as the
        # reference does (MigrationInference.py:89-99; self.times aliases the argument)
        splitFraction = splitT % 1
        splitT = int(splitT)
        if splitT - 1 > len(times):
            self.PrintError("__init__", "Invalid value for split time, cannot create Migration class instance.")
        if splitFraction != 0.0:
            t1 = splitFraction * times[splitT]
            t2 = times[splitT] - t1
            times[splitT] = t1
            times.insert(splitT + 1, t2)
            lambdas.insert(splitT + 1, lambdas[splitT])
            splitT += 1
        self.lh = list(lambdas)
        self.times = times
        self.numT = len(self.lh)
        if len(self.times) != self.numT - 1:
            print("Unexpected number of time intervals")
            sys.exit(0)
        self.discr = 1
        self.splitT = splitT
        self.mi = list(self.lh)
        self.pu = list(self.lh)
        self.SetModel(mi, pu)
        self._data_rows = None
        self.SetJAFS(dataJAFS)
        self.JAFSsize = self.snps
        self.lc = [[1, 1] for _ in range(self.numT)]
        self.M = None
        self.integralP = None
        self.P0 = None
        self.P1 = None
        self.JAFS = None
        self.Pr = None
        self.llh = None
        if self.debug:
            print("MigrationInference class initialized. Class size", self.numT)

    # ------------------------------------------------------------------ reference-compatible API
    def PrintError(self, func, text):
        func = func + "():"
        print("MigrationInference class error in function", func, text)
        sys.exit(0)

    def SetJAFS(self, dataJAFS, normalize=False):
        """MigrationInference.SetJAFS (MigrationInference.py:202-227)."""
        if len(dataJAFS) != 8:
            self.PrintError("SetJAFS", "Unexpected data SFS.")
        self.snps = sum(dataJAFS[1:])
        self.dataJAFS = dataJAFS[1:]
        if normalize:
            for i in range(1, len(self.dataJAFS)):
                self.dataJAFS[i] = self.dataJAFS[i] / self.snps * self.JAFSsize
            self.snps = self.JAFSsize
            print("True SFS size = ", self.JAFSsize, " bootstrap SFS size:", sum(self.dataJAFS))
        row = [float(dataJAFS[0])] + [float(v) for v in self.dataJAFS]
        self._set_rows(np.array([row]))
        self.llh_const = float(self._llh_consts[0])

    def SetJAFSBatch(self, rows):
        """Score every evaluation against all rows[R][8] = [total, 7 counts] (row 0 plays the role of
        dataJAFS; e.g. the output of utils/generateJSFS_bs.py: row 0 data, rows 1.. bootstrap replicates)."""
        rows = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
        self.snps = float(rows[0, 1:].sum())
        self.dataJAFS = [float(v) for v in rows[0, 1:]]
        self._set_rows(rows)
        self.llh_const = float(self._llh_consts[0])

    def _set_rows(self, rows):
        from .engine import llh_constants
        self._data_rows = np.array(rows, dtype=np.float64)
        self._llh_consts = llh_constants(self._data_rows, self.unfolded)
        self._data_dirty = True

    def SetModel(self, mis, pus):
        """MigrationInference.SetModel (MigrationInference.py:229-289): same validation, same messages."""
        self.optMis = []
        self.optPus = []
        for i in range(len(self.mi)):
            self.mi[i] = [None, None]
        for i in range(len(self.pu)):
            self.pu[i] = [None, None]
        self._bands, self._pulses = [], []
        for el in mis:
            popInd = int(el[0]) - 1
            if popInd != 0 and popInd != 1:
                self.PrintError("SetModel", "Population index should be 1 or 2.")
            migStart = int(el[1])
            if migStart < self.sampleDate:
                self.PrintError("SetModel", "Migration start (" + str(migStart) + ") should be larger than or equal to sample date (" + str(self.sampleDate) + ").")
            migEnd = int(el[2])
            if migEnd <= migStart:
                self.PrintError("SetModel", "Migration start (" + str(migStart) + ") should be strictly less than migration end (" + str(migEnd) + ").")
            migVal = float(el[3])
            migOpt = int(el[4])
            for i in range(migStart, migEnd):
                if self.mi[i][popInd] is not None:
                    self.PrintError("SetModel", "Migration rate intervals should not overlap.")
                self.mi[i][popInd] = migVal
            if migOpt == 1:
                self.optMis.append([popInd, migStart, migEnd, migVal])
            self._bands.append([popInd, migStart, migEnd, migVal, migOpt == 1])
        for el in pus:
            popInd = int(el[0]) - 1
            if popInd != 0 and popInd != 1:
                self.PrintError("SetModel", "Population index should be 1 or 2.")
            puTime = int(el[1])
            if puTime < self.sampleDate:
                self.PrintError("SetModel", "Pulse migration time (" + str(puTime) + ") should be larger than or equal to sample date (" + str(self.sampleDate) + ").")
            puVal = float(el[2])
            if puVal < 0 or puVal > 1:
                self.PrintError("SetModel", "Pulse migration rate should be between 0 and 1.")
            puOpt = int(el[3])
            if self.pu[puTime][0] is not None or self.pu[puTime][1] is not None:
                self.PrintError("SetModel", "Current version allows only single-direction pulse migration at a time.")
            self.pu[puTime][popInd] = puVal
            if puOpt == 1:
                self.optPus.append([popInd, puTime, puVal])
            self._pulses.append([popInd, puTime, puVal, puOpt == 1])
        for arr in (self.mi, self.pu):
            for i in range(len(arr)):
                for k in (0, 1):
                    if arr[i][k] is None:
                        arr[i][k] = 0.0
        self.optMisSize = len(self.optMis)
        self.optPusSize = len(self.optPus)
        self._registered = False

    def MapParameters(self, params):
        """MigrationInference.MapParameters (MigrationInference.py:291-298)."""
        if len(params) != self.optMisSize + self.optPusSize:
            self.PrintError("MapParameters", "Incorrect number of parameters.")
        for i in range(self.optMisSize):
            for j in range(self.optMis[i][1], self.optMis[i][2]):
                self.mi[j][self.optMis[i][0]] = params[i]
        for i in range(self.optPusSize):
            self.pu[self.optPus[i][1]][self.optPus[i][0]] = params[self.optMisSize + i]

    # ------------------------------------------------------------------ device plumbing
    def _flags(self):
        f = 0
        if self.correct:
            f |= _lib.FLAG_CORRECT
        if self.cpfit:
            f |= _lib.FLAG_CPFIT
        if self.smooth:
            f |= _lib.FLAG_SMOOTH
        if self.unfolded:
            f |= _lib.FLAG_UNFOLDED
        return f

    def register_into(self, eng, grid_id=None):
        """Register this object's grid (unless grid_id is given) and model layout in an Engine; returns
        (grid_id, model_id).  Used by the object itself and by misti_b200.sweep for shared engines."""
        gid = eng.add_grid(self.times, self.lh) if grid_id is None else grid_id
        bands, pulses, k = [], [], 0
        # optimiser index order = MapParameters order: optimised bands first, then optimised pulses
        for pop, a, b, val, opt in self._bands:
            bands.append((pop, a, min(b, self.numT), val, k if opt else -1))
            k += 1 if opt else 0
        for pop, t, val, opt in self._pulses:
            pulses.append((pop, t, val, k if opt else -1))
            k += 1 if opt else 0
        return gid, eng.add_model(gid, self.splitT, self.sampleDate, bands, pulses)

    def _sync_engine(self):
        if self._engine is None:
            self._engine = Engine(self._device)
        eng = self._engine
        if not self._registered:
            if not self._owns_engine:
                raise RuntimeError("a shared Engine must be populated by its owner (use misti_b200.sweep)")
            eng.clear_models()
            _, self._model_id = self.register_into(eng)
            self._registered = True
            self._data_dirty = True
        if self._data_dirty:
            eng.set_data(self._data_rows, self.unfolded, self._llh_consts)
            self._data_dirty = False
        return eng

    def _current_params(self):
        """The optimiser vector implied by the current self.mi / self.pu (after MapParameters)."""
        p = [self.mi[b[1]][b[0]] for b in self.optMis]
        p += [self.pu[q[1]][q[0]] for q in self.optPus]
        return p

    def _evaluate(self, params, want, lc_inject=None):
        eng = self._sync_engine()
        P = self.optMisSize + self.optPusSize
        params = np.asarray(params, dtype=np.float64).reshape(-1, P) if P else np.zeros((len(params), 0))
        return eng.evaluate(params, model=self._model_id, flags=self._flags(), mixtureTH=self.mixtureTH,
                            lc_inject=lc_inject, want=want)

    def _store_chain(self, out):
        lc = out["lc"][0]
        self.lc = [[float(lc[t, 0]), float(lc[t, 1])] for t in range(self.numT)]
        n_pr = min(self.splitT, self.numT) + 1
        pr = out["pr"][0]
        self.Pr = [[[float(pr[t, s, 0]), float(pr[t, s, 1])] for s in range(3)] for t in range(n_pr)]

    # ------------------------------------------------------------------ the hot path
    def CorrectLambdas(self):
        """MigrationInference.CorrectLambdas (MigrationInference.py:305-378) + Smooth, on the device.
        Fills self.lc / self.Pr; returns False where the reference does."""
        MigrationInference.CORRECTION_CALLED += 1
        self._note_cl_mu()
        out = self._evaluate([self._current_params()], want=("lc", "pr", "status"))
        st = int(out["status"][0])
        self._store_chain(out)
        if st == _lib.CORRECTION_FAILED:
            MigrationInference.CORRECTION_FAILED += 1
        return st not in (_lib.CORRECTION_FAILED, _lib.NEGATIVE_PARAM)

    def _note_cl_mu(self):
        """CorrectLambdas leaves the migration rates of the last interval it visits in the CorrectLambda helper
        (cl.SetMu, MigrationInference.py:324); CoalescentRates later uses exactly those for every interval."""
        if self.splitT > 0:
            t = min(self.splitT, self.numT) - 1
            self._cl_mu = [float(self.mi[t][0]), float(self.mi[t][1])]

    def CoalescentRates(self):
        """MigrationInference.CoalescentRates (MigrationInference.py:542-564): the forward map, on the device -- the
        object's rates self.lh are taken as the TRUE model rates (copied to self.lc) and replaced, before the split, by the
        rates PSMC would see; self.Pr gets the trajectory of the 3-state chains.  Like the reference the method uses, for
        EVERY interval, the migration rates the preceding CorrectLambdas / JAFSLikelihood call left in the helper (those of
        the last interval before the split), and fails with AttributeError when there was no such call (TestModel.py:96,120
        calls JAFSLikelihood first)."""
        if getattr(self, "_cl_mu", None) is None:
            raise AttributeError("'CorrectLambda' object has no attribute 'mu'")
        eng = self._sync_engine()
        lh, pr = eng.coalescent_rates(self._model_id, self._current_params(), self._cl_mu)
        self.lc = [[float(v[0]), float(v[1])] for v in self.lh]
        for t in range(min(self.splitT, self.numT)):
            self.lh[t][0], self.lh[t][1] = float(lh[t, 0]), float(lh[t, 1])
        self.Pr = [[[float(pr[t, s, 0]), float(pr[t, s, 1])] for s in range(3)] for t in range(pr.shape[0])]
        self._registered = False  # the grid registered in the engine still holds the true rates

    def JAFSpectrum(self):
        """MigrationInference.JAFSpectrum (MigrationInference.py:467-506) for the CURRENT self.lc
        (injected into the device evaluation); sets self.JAFS to the unnormalised spectrum."""
        inj = np.zeros((1, self._sync_engine().numT_max, 2))
        inj[0, :self.numT, :] = np.asarray(self.lc, dtype=np.float64)
        out = self._evaluate([self._current_params()], want=("jafs_raw", "status"), lc_inject=inj)
        st = int(out["status"][0])
        if st == _lib.INFINITE_COAL_TIME:
            print("Infinite coalescent time. No migration.")
            sys.exit(0)
        self.JAFS = [float(v) for v in out["jafs_raw"][0]]
        return self.JAFS

    def JAFSLikelihood(self, mu):
        """MigrationInference.JAFSLikelihood (MigrationInference.py:566-614): one evaluation on the device."""
        MigrationInference.COUNT_LLH += 1
        self.llh = -10 ** 9
        for v in mu:
            if v < 0:
                print("Hit negative value of migration rate")
                return -np.inf
        self.MapParameters(mu)
        MigrationInference.CORRECTION_CALLED += 1
        self._note_cl_mu()
        out = self._evaluate([list(mu)], want=("jafs", "lc", "pr", "status"))
        st = int(out["status"][0])
        self._store_chain(out)
        if st == _lib.CORRECTION_FAILED or st == _lib.NEGATIVE_PARAM:
            MigrationInference.CORRECTION_FAILED += 1
            print("Lambda correction failed")
            return -np.inf
        if st == _lib.INFINITE_COAL_TIME:
            print("Infinite coalescent time. No migration.")
            sys.exit(0)
        if self.enableOutput:
            print("JAFSLikelihood():   initial values of lambdas are ", self.lh)
            print("JAFSLikelihood(): corrected values of lambdas are ", self.lc)
        self.JAFS = [float(v) for v in out["jafs"][0]]
        llh = float(out["llh"][0, 0])
        self.llh = llh
        return llh

    def JAFSLikelihoodBatch(self, params, return_all=False):
        """Batched objective: params[B, P] -> llh[B] against dataJAFS, or llh[B, R] against all rows set
        with SetJAFSBatch.  -inf where JAFSLikelihood would return -inf.  With return_all=True returns the
        engine's dict (llh, jafs, status, nfev, terms)."""
        P = self.optMisSize + self.optPusSize
        params = np.asarray(params, dtype=np.float64).reshape(-1, P) if P else np.zeros((len(params), 0))
        B = params.shape[0]
        MigrationInference.COUNT_LLH += B
        MigrationInference.CORRECTION_CALLED += B
        out = self._evaluate(params, want=("jafs", "status", "nfev", "terms"))
        MigrationInference.CORRECTION_FAILED += int(np.count_nonzero(out["status"] == _lib.CORRECTION_FAILED))
        if return_all:
            return out
        llh = out["llh"]
        return llh[:, 0] if llh.shape[1] == 1 else llh

    def ExpectedJAFSBatch(self, params):
        """params[B, P] -> normalised expected JSFS [B, 7] (NaN rows where the evaluation failed)."""
        return self.JAFSLikelihoodBatch(params, return_all=True)["jafs"]

    def MaximumLLHFunction(self):
        """MigrationInference.MaximumLLHFunction (MigrationInference.py:696-711): the data spectrum scored
        against itself -- the likelihood tail kernel run on the normalised data."""
        eng = self._sync_engine()
        return float(eng.score_spectra([self.dataJAFS])[0, 0])

    def ObjectiveFunction(self, mu):
        res = -self.JAFSLikelihood(mu)
        print(mu, res)
        return res

    def Solve(self, tol=1e-4, globalOpt=False):
        """MigrationInference.Solve (MigrationInference.py:718-733): the same scipy drivers around the
        device objective, so the simplex sequence is the reference's."""
        from scipy import optimize
        if self.optMisSize + self.optPusSize > 0:
            init = [val[3] for val in self.optMis] + [val[2] for val in self.optPus]
            if globalOpt:
                res = optimize.basinhopping(self.ObjectiveFunction, init, T=0.5, minimizer_kwargs=dict(method='Nelder-Mead'))
            else:
                res = optimize.minimize(self.ObjectiveFunction, init, method='Nelder-Mead',
                                        options={'xatol': tol, 'fatol': tol, 'maxiter': 1000, 'disp': True})
            return [res.x, -res.fun]
        return [[], self.JAFSLikelihood([])]

    def SolveBatch(self, tol=1e-4, x0=None, globalOpt=False, niter=100, seeds=None):
        """Addition: Solve() for EVERY data row set with SetJAFSBatch at once -- one Nelder-Mead fit per
        row (e.g. per bootstrap replicate), all taken through scipy's decisions on the device (Engine.nelder_mead: no
        host round trip per step, nothing printed per evaluation).  x0: start vector(s) [P] or [R, P]; default = the
        model's initial values as in Solve.  Returns (x [R, P], llh [R], info) with info = the optimiser's dict
        (nit, nfev, status, success as scipy counts them).  Row r of the result equals what Solve(tol) returns for a
        model whose data are row r.
        globalOpt: Solve(globalOpt=True) per row instead -- one basin-hopping walker per row on the device
        (Engine.basinhopping: T = 0.5, scipy's defaults otherwise, `niter` hops; the reference passes no seed, here walker r
        draws the numbers of numpy.random.default_rng(seeds[r]), default seeds 0, 1, ...)."""
        eng = self._sync_engine()
        P = self.optMisSize + self.optPusSize
        R = len(self._data_rows)
        if P == 0:
            out = self._evaluate([[]], want=("status",))
            MigrationInference.COUNT_LLH += 1
            return np.zeros((R, 0)), out["llh"][0].copy(), {"nfev": np.ones(R, dtype=np.int64)}
        init = [val[3] for val in self.optMis] + [val[2] for val in self.optPus] if x0 is None else x0
        X0 = np.broadcast_to(np.asarray(init, dtype=np.float64).reshape(-1, P), (R, P)).copy()
        if globalOpt:
            r = eng.basinhopping(X0, np.full(R, self._model_id, dtype=np.int32), np.arange(R, dtype=np.int32),
                                 seeds=list(range(R)) if seeds is None else list(seeds), flags=self._flags(), mixtureTH=self.mixtureTH,
                                 niter=niter, T=0.5)
            MigrationInference.COUNT_LLH += r["evaluations"]
            MigrationInference.CORRECTION_CALLED += r["evaluations"]
            return r["x"], -r["fun"], r
        r = eng.nelder_mead(X0, np.full(R, self._model_id, dtype=np.int32), np.arange(R, dtype=np.int32), flags=self._flags(),
                            mixtureTH=self.mixtureTH, xatol=tol, fatol=tol, maxiter=1000)
        MigrationInference.COUNT_LLH += r["evaluations"]
        MigrationInference.CORRECTION_CALLED += r["evaluations"]
        return r["x"], -r["fun"], r

    @staticmethod
    def Report():
        print("Total number of likelihood function calls is", MigrationInference.COUNT_LLH)
        print("Lambda correction called", MigrationInference.CORRECTION_CALLED, "times.")
        print("Lambda correction failed", MigrationInference.CORRECTION_FAILED, "times.")
