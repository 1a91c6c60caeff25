"""Sweep: every (data row x split time x band layout) fit of a study in ONE process on ONE engine.

Replaces the reference's workflow of one MiSTI.py process per (bootstrap row, split time) under bash loops /
GNU parallel (README.md:110-117, test.bs/san_sar.bs.no.mig.sh:29-36, test.bs/din_sar.bs.sh:29-38): the PSMC
grid is registered once, each candidate model (split time, -mi bands, -pu pulses) once, the data and bootstrap
rows once, and all fits advance in lock step (misti_b200.optim) with one device launch per optimiser step.
Result lines are formatted like MiSTI.py:240, which the reference's downstream tooling greps.
"""
from math import ceil

import numpy as np

from . import _lib
from .engine import Engine, llh_constants
from .inference import MigrationInference
from .optim import basinhopping_batch, nelder_mead_batch


def split_time_interval(best_split_per_row, level=0.975):
    """The reduction of the reference's bootstrap notebook (test.bs/bs_conf_int.ipynb: conf_int_bs): row 0 is the data, rows
    1.. are bootstrap replicates, each with the split time of its highest likelihood; returns dict(interval = Student-t interval
    of the replicates' mean at `level` (the notebook's scipy.stats.t.interval(0.975, n - 1, loc = mean, scale = sem)),
    from_data = row 0's split time, histogram = {split time: replicates})."""
    from collections import Counter
    from scipy import stats
    best = [float(v) for v in best_split_per_row]
    a = np.array(best[1:], dtype=np.float64)
    if a.size < 2:
        raise ValueError("need the data row and at least two bootstrap replicates")
    sem = stats.sem(a)
    lo, hi = stats.t.interval(level, len(a) - 1, loc=np.mean(a), scale=sem) if sem > 0 else (float(np.mean(a)), float(np.mean(a)))
    return {"interval": (float(lo), float(hi)), "from_data": best[0], "histogram": dict(Counter(best[1:]))}


class Sweep:
    def __init__(self, times, lambdas, rows, unfolded=False, cpfit=False, smooth=True, trueEPS=False, sampleDate=0,
                 mixtureTH=0.0, engine=None, device=0):
        """times / lambdas: the merged PSMC grid (InputData.times / .lambdas); rows: [R][8] data + bootstrap rows."""
        self.times = [float(v) for v in times]
        self.lambdas = [[float(v[0]), float(v[1])] for v in lambdas]
        self.rows = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
        self.kw = dict(unfolded=unfolded, cpfit=cpfit, smooth=smooth, trueEPS=trueEPS, sampleDate=sampleDate, mixtureTH=mixtureTH)
        self.engine = engine if engine is not None else Engine(device)
        self.engine.clear_models()
        self.engine.set_data(self.rows, unfolded, llh_constants(self.rows, unfolded))
        self.flags = (0 if trueEPS else _lib.FLAG_CORRECT) | (_lib.FLAG_CPFIT if cpfit else 0) | \
                     (_lib.FLAG_SMOOTH if smooth else 0) | (_lib.FLAG_UNFOLDED if unfolded else 0)
        self.mixtureTH = float(mixtureTH)
        self.models = []     # per model: dict(id, splitT (as given), mi, pu, init, n_params, host)
        self._base_grid = None
        self._mid_arr = None

    def add_model(self, splitT, mi=(), pu=()):
        """Same arguments as the MiSTI.py command line: split time (may be fractional), -mi 5-tuples, -pu 4-tuples."""
        times, lam = list(self.times), [list(v) for v in self.lambdas]
        host = MigrationInference(times, lam, list(self.rows[0]), splitT, [list(m) for m in mi], [list(p) for p in pu],
                                  engine=self.engine, **self.kw)
        frac = float(splitT) % 1 != 0.0
        if not frac and self._base_grid is None:
            self._base_grid = self.engine.add_grid(self.times, self.lambdas)
        gid, mid = host.register_into(self.engine, None if frac else self._base_grid)
        init = [b[3] for b in host.optMis] + [q[2] for q in host.optPus]
        self.models.append(dict(id=mid, splitT=splitT, mi=[list(m) for m in mi], pu=[list(p) for p in pu], init=init,
                                n_params=len(init), host=host))
        return len(self.models) - 1

    def _model_ids(self):
        """engine model id of every model of the sweep, as an array (looked up per item of every batch)"""
        if self._mid_arr is None or len(self._mid_arr) != len(self.models):
            self._mid_arr = np.array([m["id"] for m in self.models], dtype=np.int32)
        return self._mid_arr

    # -- evaluation of arbitrary (model, params, row) triples ----------------------------------------
    def evaluate(self, model_idx, params, row_idx):
        model_idx = np.asarray(model_idx, dtype=np.int64).reshape(-1)
        K = model_idx.shape[0]
        P = max([m["n_params"] for m in self.models] + [0])
        X = np.zeros((K, P))
        params = np.asarray(params, dtype=np.float64).reshape(K, -1) if K else np.zeros((0, P))
        X[:, :params.shape[1]] = params
        mids = self._model_ids()[model_idx]
        out = self.engine.evaluate(X, model_ids=mids, flags=self.flags, mixtureTH=self.mixtureTH, want=("status",),
                                   row_ids=np.asarray(row_idx, dtype=np.int32))
        MigrationInference.COUNT_LLH += K
        MigrationInference.CORRECTION_CALLED += K
        MigrationInference.CORRECTION_FAILED += int(np.count_nonzero(out["status"] == _lib.CORRECTION_FAILED))
        return out["llh"][:, 0], out["status"]

    def evaluate_grid(self, params_per_model=None):
        """llh[model, row] at fixed parameters (default: each model's initial values) -- config 5 without migration."""
        M, R = len(self.models), self.rows.shape[0]
        P = max([m["n_params"] for m in self.models] + [0])
        X = np.zeros((M, P))
        for i, m in enumerate(self.models):
            v = m["init"] if params_per_model is None else params_per_model[i]
            X[i, :len(v)] = v
        mids = np.array([m["id"] for m in self.models], dtype=np.int32)
        out = self.engine.evaluate(X, model_ids=mids, flags=self.flags, mixtureTH=self.mixtureTH, want=("status",))
        MigrationInference.COUNT_LLH += M
        return out["llh"]

    def argmax_split(self, params_per_model=None):
        """Per data row (bootstrap replicate) the model -- split time -- of the highest likelihood at fixed parameters, and
        that likelihood: the reduction of the reference's bootstrap notebook (test.bs/bs_conf_int.ipynb: per replicate the
        arg-max over `st` of the result lines), taken on the device, so that M x R likelihoods come back as 2 R numbers.
        Returns dict(model [R] index into self.models, splitT [R], llh [R])."""
        M = len(self.models)
        P = max([m["n_params"] for m in self.models] + [0])
        X = np.zeros((M, P))
        for i, m in enumerate(self.models):
            v = m["init"] if params_per_model is None else params_per_model[i]
            X[i, :len(v)] = v
        out = self.engine.evaluate(X, model_ids=self._model_ids(), flags=self.flags, mixtureTH=self.mixtureTH, want=(), row_best="only")
        MigrationInference.COUNT_LLH += M
        idx = out["row_best_item"].astype(np.int64)
        return {"model": idx, "splitT": np.array([self.models[i]["splitT"] for i in idx]), "llh": out["row_best_llh"]}

    def split_time_confidence(self, res=None, level=0.975):
        """Confidence interval of the split time over the bootstrap rows, as the reference's notebook computes it from one
        result line per (row, split time): per row the split time of the highest likelihood -- at fixed parameters reduced on
        the device (argmax_split), or taken from a solve() result `res` (fitted likelihoods) -- then split_time_interval."""
        R = self.rows.shape[0]
        if res is None:
            best = self.argmax_split()["splitT"]
        else:
            best = np.empty(R)
            for r in range(R):
                k = np.nonzero(np.asarray(res["row"]) == r)[0]
                best[r] = self.models[int(np.asarray(res["model"])[k[np.argmax(np.asarray(res["llh"])[k])]])]["splitT"]
        return split_time_interval(best, level)

    # -- fits ------------------------------------------------------------------------------------------
    # up to this many points per round the fits are stepped on the device (Engine.nelder_mead / Engine.basinhopping: no host
    # round trip per step, only the simplices still running are packed into a round); beyond, the host driver takes over
    DEVICE_NM_MAX_POINTS = (1 << 18) - 4096  # misti_fit: every simplex's own slots + the shared look-ahead region fit one launch

    def solve(self, pairs=None, tol=1e-4, globalOpt=False, niter=100, seed=0, speculative=True, on_device="auto"):
        """Fit every (model, row) pair (default: all).  Nelder-Mead with xatol = fatol = tol, maxiter = 1000 as
        MigrationInference.Solve; globalOpt = basin-hopping with T = 0.5 as the reference calls it (scipy-default
        inner tolerances), seeded per pair.  Returns a dict of arrays indexed by pair.
        on_device: take the Nelder-Mead steps on the device (True / False / "auto" = by the size of a round); the
        decisions, iterates and counts are the same either way."""
        M, R = len(self.models), self.rows.shape[0]
        if pairs is None:
            pairs = [(m, r) for r in range(R) for m in range(M)]
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        K = pairs.shape[0]
        Pmax = max([m["n_params"] for m in self.models] + [0])
        res = dict(model=pairs[:, 0], row=pairs[:, 1], x=np.full((K, Pmax), np.nan), llh=np.full(K, np.nan),
                   nfev=np.zeros(K, dtype=np.int64), nit=np.zeros(K, dtype=np.int64), success=np.zeros(K, dtype=bool),
                   evaluations=0, launches=0)
        nparams = np.array([self.models[m]["n_params"] for m in pairs[:, 0]])
        fixed = np.nonzero(nparams == 0)[0]
        if fixed.size:
            llh, _ = self.evaluate(pairs[fixed, 0], np.zeros((fixed.size, 0)), pairs[fixed, 1])
            res["llh"][fixed] = llh
            res["nfev"][fixed] = 1
            res["success"][fixed] = np.isfinite(llh)
            res["evaluations"] += fixed.size
            res["launches"] += 1
        for P in sorted(set(nparams[nparams > 0].tolist())):
            sel = np.nonzero(nparams == P)[0]
            self.engine.reserve(sel.size * max(4, P + 1), Pmax)  # a step submits four candidates (or a shrink) per simplex
            x0 = np.array([self.models[m]["init"] for m in pairs[sel, 0]], dtype=np.float64)

            def fun(X, who, sel=sel):
                llh, _ = self.evaluate(pairs[sel[who], 0], X, pairs[sel[who], 1])
                return -llh
            dev = on_device is True or (on_device == "auto" and speculative and sel.size * max(4, P + 1) <= self.DEVICE_NM_MAX_POINTS)
            mids = self._model_ids()[pairs[sel, 0]]

            def device_nm(xs, mids=mids, rows=pairs[sel, 1], **kw):
                r = self.engine.nelder_mead(xs, mids, rows, flags=self.flags, mixtureTH=self.mixtureTH, **kw)
                MigrationInference.COUNT_LLH += r["evaluations"]
                MigrationInference.CORRECTION_CALLED += r["evaluations"]
                return r
            if globalOpt and dev:  # walkers on the device: no barrier between the hops of different walkers
                r = self.engine.basinhopping(x0, mids, pairs[sel, 1], seeds=[seed + int(k) for k in sel], flags=self.flags,
                                             mixtureTH=self.mixtureTH, niter=niter, T=0.5)
                MigrationInference.COUNT_LLH += r["evaluations"]
                MigrationInference.CORRECTION_CALLED += r["evaluations"]
            elif globalOpt:
                r = basinhopping_batch(fun, x0, niter=niter, T=0.5, seeds=[seed + int(k) for k in sel], speculative=speculative)
            else:
                if dev:
                    r = device_nm(x0, xatol=tol, fatol=tol, maxiter=1000)
                else:
                    r = nelder_mead_batch(fun, x0, xatol=tol, fatol=tol, maxiter=1000, speculative=speculative)
                res["nit"][sel] = r["nit"]
            res["x"][sel, :P] = r["x"]
            res["llh"][sel] = -r["fun"]
            res["nfev"][sel] = r["nfev"]
            res["success"][sel] = r["success"]
            res["evaluations"] += r["evaluations"]
            res["launches"] += r["launches"]
        return res

    def solve_distributed(self, pairs=None, group=None, device=None, **kw):
        """solve() with the (model, row) pairs dealt over the ranks of the torch.distributed job (one process per GPU, every
        rank holding the same Sweep): each rank fits its share, one all-gather brings x, llh and the counts of every pair
        to every rank (misti_b200.parallel.solve_sharded).  Without a process group it is solve()."""
        from .parallel import solve_sharded
        M, R = len(self.models), self.rows.shape[0]
        if pairs is None:
            pairs = [(m, r) for r in range(R) for m in range(M)]
        return solve_sharded(lambda shard: self.solve(pairs=shard, **kw), pairs, group=group, device=device)

    def result_line(self, res, k, scaleTime=1.0, bs_id=None):
        """The reference's result line (MiSTI.py:240) for pair k of a solve() result."""
        m = self.models[int(res["model"][k])]
        fixed = [float(el[3]) for el in m["mi"] if int(el[4]) == 0]
        x = [v for v in res["x"][k][:m["n_params"]]]
        fs = "fixed = [" + ", ".join(str(v) for v in fixed) + "]" if fixed else ""
        os_ = "optim = [" + ", ".join(str(v) for v in x) + "]" if x else ""
        mig = fs + "\t" + os_ if fs and os_ else fs + os_
        # MiSTI.py:240 sums the caller's list AFTER the constructor has cut the split interval in two (fractional split
        # times): of that interval only the fraction counts -- the model's own (cut) grid, not the sweep's
        t = sum(m["host"].times[0:ceil(float(m["splitT"]))]) * scaleTime
        row = int(res["row"][k]) if bs_id is None else bs_id
        return "bs_id = %s \tsplitT = %s \ttime = %s \tmigration rates %s \tllh = %s" % (row, m["splitT"], t, mig, res["llh"][k])
