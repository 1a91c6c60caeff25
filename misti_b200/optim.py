"""Batched optimisers around the batched objective: many independent Nelder-Mead simplices / basin-hopping
walkers advanced in LOCK STEP, one device launch per step for all of them.

The reference drives the objective serially from ``scipy.optimize.minimize(method='Nelder-Mead')`` or
``scipy.optimize.basinhopping`` (MigrationInference.Solve, MigrationInference.py:718-733), one process per
(bootstrap row, split time) in the test.bs scripts.  Here every simplex takes exactly the decisions scipy
1.18.1 takes (``scipy/optimize/_optimize.py:_minimize_neldermead``: rho=1, chi=2, psi=0.5, sigma=0.5, initial
simplex x0*(1.05) or 0.00025, strict / non-strict comparisons, stable re-ordering, termination
max|x_i - x_0| <= xatol and max|f_i - f_0| <= fatol), so given the same objective values the result, the
iteration count and scipy's function-evaluation count are identical -- the candidates of a step (reflection,
expansion, outside and inside contraction) are merely evaluated together, speculatively, instead of one after
the other.  Basin-hopping follows ``scipy/optimize/_basinhopping.py`` (uniform displacement, adaptive step size
every `interval` steps, Metropolis test at temperature T, best-so-far storage) with one numpy Generator per
walker, so a single walker with seed s reproduces ``basinhopping(..., rng=s)``.

`fun(X[K, N], owner[K]) -> f[K]` evaluates K points at once; owner[k] is the index of the simplex / walker the
point belongs to (so the caller can attach a model id or a data row to it).  NaN objective values are treated
as +inf.  Pure host-side control logic (numpy); all arithmetic of the objective is on the device.
"""
import math

import numpy as np

RHO, CHI, PSI, SIGMA = 1.0, 2.0, 0.5, 0.5
NONZDELT, ZDELT = 0.05, 0.00025


def _clean(f):
    f = np.asarray(f, dtype=np.float64).reshape(-1).copy()
    f[np.isnan(f)] = np.inf
    return f


def _sort(sim, fsim):
    ind = np.argsort(fsim, axis=1, kind="stable")
    return np.take_along_axis(sim, ind[:, :, None], axis=1), np.take_along_axis(fsim, ind, axis=1)


def initial_simplex(x0):
    """scipy's default simplex (_optimize.py:775-801) for every row of x0[S, N] -> [S, N+1, N]."""
    x0 = np.asarray(x0, dtype=np.float64)
    S, N = x0.shape
    sim = np.empty((S, N + 1, N))
    sim[:, 0] = x0
    for k in range(N):
        y = x0.copy()
        y[:, k] = np.where(y[:, k] != 0, (1 + NONZDELT) * y[:, k], ZDELT)
        sim[:, k + 1] = y
    return sim


def nelder_mead_batch(fun, x0, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, speculative=True, owners=None):
    """Minimise S independent objectives in lock step.  x0: [S, N].  Returns a dict of arrays:
    x [S, N], fun [S], nit [S], nfev [S] (scipy's count), status [S] (0 converged, 1 maxfev, 2 maxiter),
    success [S], evaluations (points actually sent to `fun`), launches (calls of `fun`)."""
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.ndim == 1:
        x0 = x0.reshape(1, -1)
    S, N = x0.shape
    owners = np.arange(S) if owners is None else np.asarray(owners)
    if maxiter is None and maxfev is None:
        maxiter, maxfev = N * 200, N * 200
    elif maxiter is None:
        maxiter = N * 200 if maxfev == np.inf else np.inf
    elif maxfev is None:
        maxfev = N * 200 if maxiter == np.inf else np.inf
    sim = initial_simplex(x0)
    evaluations, launches = 0, 0

    def call(X, who):
        nonlocal evaluations, launches
        evaluations += len(X)
        launches += 1
        return _clean(fun(np.ascontiguousarray(X), owners[who]))

    fsim = call(sim.reshape(-1, N), np.repeat(np.arange(S), N + 1)).reshape(S, N + 1)
    fcalls = np.full(S, N + 1, dtype=np.int64)
    sim, fsim = _sort(sim, fsim)
    iterations = np.ones(S, dtype=np.int64)
    status = np.full(S, -1, dtype=np.int64)
    active = np.ones(S, dtype=bool)
    while True:
        budget = (fcalls < maxfev) & (iterations < maxiter)
        with np.errstate(invalid="ignore"):
            conv = (np.max(np.abs(sim[:, 1:] - sim[:, :1]).reshape(S, -1), axis=1) <= xatol) & \
                   (np.max(np.abs(fsim[:, :1] - fsim[:, 1:]), axis=1) <= fatol)
        done_now = active & (~budget | conv)
        status[done_now & budget & conv] = 0
        status[done_now & ~budget & (fcalls >= maxfev)] = 1
        status[done_now & ~budget & (fcalls < maxfev)] = 2
        active &= ~done_now
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        A = idx.size
        s, f = sim[idx], fsim[idx]
        xbar = s[:, 0].copy()
        for j in range(1, N):  # np.add.reduce(sim[:-1], 0): sequential row sum
            xbar = xbar + s[:, j]
        xbar = xbar / N
        last = s[:, -1]
        xr = (1 + RHO) * xbar - RHO * last
        xe = (1 + RHO * CHI) * xbar - RHO * CHI * last
        xc = (1 + PSI * RHO) * xbar - PSI * RHO * last
        xcc = (1 - PSI) * xbar + PSI * last
        if speculative:
            fall = call(np.concatenate([xr, xe, xc, xcc]), np.tile(idx, 4)).reshape(4, A)
            fxr, fxe, fxc, fxcc = fall
        else:
            fxr = call(xr, idx)
            want_e = fxr < f[:, 0]
            want_c = ~want_e & ~(fxr < f[:, -2]) & (fxr < f[:, -1])
            want_cc = ~want_e & ~(fxr < f[:, -2]) & ~(fxr < f[:, -1])
            second = np.where(want_e[:, None], xe, np.where(want_c[:, None], xc, xcc))
            need = want_e | want_c | want_cc
            f2 = np.full(A, np.inf)
            if need.any():
                f2[need] = call(second[need], idx[need])
            fxe, fxc, fxcc = f2, f2, f2
        # scipy's decision tree (_optimize.py:846-896), vectorised
        better_than_best = fxr < f[:, 0]
        take_e = better_than_best & (fxe < fxr)
        take_r = (better_than_best & ~(fxe < fxr)) | (~better_than_best & (fxr < f[:, -2]))
        contract = ~better_than_best & ~(fxr < f[:, -2])
        outside = contract & (fxr < f[:, -1])
        inside = contract & ~(fxr < f[:, -1])
        take_c = outside & (fxc <= fxr)
        take_cc = inside & (fxcc < f[:, -1])
        shrink = (outside & ~take_c) | (inside & ~take_cc)
        new_x = np.where(take_e[:, None], xe, np.where(take_r[:, None], xr, np.where(take_c[:, None], xc, xcc)))
        new_f = np.where(take_e, fxe, np.where(take_r, fxr, np.where(take_c, fxc, fxcc)))
        replace = take_e | take_r | take_c | take_cc
        s[replace, -1] = new_x[replace]
        f[replace, -1] = new_f[replace]
        fcalls[idx] += 1 + (better_than_best | contract).astype(np.int64)
        if shrink.any():
            sh = np.nonzero(shrink)[0]
            s[sh, 1:] = s[sh, :1] + SIGMA * (s[sh, 1:] - s[sh, :1])
            fs = call(s[sh, 1:].reshape(-1, N), np.repeat(idx[sh], N)).reshape(len(sh), N)
            f[sh, 1:] = fs
            fcalls[idx[sh]] += N
        s, f = _sort(s, f)
        sim[idx], fsim[idx] = s, f
        iterations[idx] += 1
    return {"x": sim[:, 0].copy(), "fun": fsim.min(axis=1), "nit": iterations, "nfev": fcalls, "status": status,
            "success": status == 0, "sim": sim, "fsim": fsim, "evaluations": evaluations, "launches": launches}


def basinhopping_batch(fun, x0, niter=100, T=1.0, stepsize=0.5, interval=50, target_accept_rate=0.5, stepwise_factor=0.9,
                       seeds=None, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, speculative=True):
    """W basin-hopping walkers in lock step (scipy/optimize/_basinhopping.py), local search = nelder_mead_batch.
    x0: [W, N]; seeds: one seed (or numpy Generator) per walker.  Returns dict: x [W, N], fun [W], nfev [W],
    nit, accepted [W], minimization_failures [W], evaluations, launches."""
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.ndim == 1:
        x0 = x0.reshape(1, -1)
    W, N = x0.shape
    seeds = list(range(W)) if seeds is None else list(seeds)
    rngs = [s if isinstance(s, np.random.Generator) else np.random.default_rng(s) for s in seeds]
    beta = 1.0 / T if T != 0 else float("inf")
    evaluations, launches = 0, 0

    def local(xs):
        nonlocal evaluations, launches
        r = nelder_mead_batch(fun, xs, xatol=xatol, fatol=fatol, maxiter=maxiter, maxfev=maxfev, speculative=speculative)
        evaluations += r["evaluations"]
        launches += r["launches"]
        return r

    r = local(x0)
    x, energy, ok = r["x"].copy(), r["fun"].copy(), r["success"].copy()
    best_x, best_f, best_ok = x.copy(), energy.copy(), ok.copy()
    nfev = r["nfev"].copy()
    failures = (~ok).astype(np.int64)
    step = np.full(W, float(stepsize))
    nstep = np.zeros(W, dtype=np.int64)
    naccept = np.zeros(W, dtype=np.int64)
    for _ in range(niter):
        # AdaptiveStepsize.take_step: count, adapt every `interval` steps, then displace
        nstep += 1
        adapt = nstep % interval == 0
        if adapt.any():
            rate = naccept / np.maximum(nstep, 1)
            step = np.where(adapt, np.where(rate > target_accept_rate, step / stepwise_factor, step * stepwise_factor), step)
        trial = np.stack([x[w] + rngs[w].uniform(-step[w], step[w], N) for w in range(W)])
        r = local(trial)
        nfev += r["nfev"]
        failures += (~r["success"]).astype(np.int64)
        fn, xn, okn = r["fun"], r["x"], r["success"]
        for w in range(W):  # Metropolis.accept_reject, with Python's min(0, nan) == 0 semantics
            prod = -(fn[w] - energy[w]) * beta
            wgt = math.exp(min(0, prod))
            accept = wgt >= rngs[w].uniform() and (okn[w] or not ok[w])
            if accept:
                naccept[w] += 1
                energy[w], x[w], ok[w] = fn[w], xn[w], okn[w]
                if okn[w] and (fn[w] < best_f[w] or not best_ok[w]):  # Storage.update
                    best_f[w], best_x[w], best_ok[w] = fn[w], xn[w], okn[w]
    return {"x": best_x, "fun": best_f, "success": best_ok, "nfev": nfev, "nit": niter, "accepted": naccept,
            "minimization_failures": failures, "evaluations": evaluations, "launches": launches}
