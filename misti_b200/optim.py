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
LOOKAHEAD_MAX_POINTS = 4096  # look-ahead (two Nelder-Mead iterations per call) while a call stays this small: up to here
                             # the device evaluates a batch as fast as a single point (profiles/r01_latency.json)
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


def _candidates(s):
    """reflection, expansion, outside and inside contraction points of sorted simplices s [A, N+1, N] (_optimize.py:846-874)"""
    N = s.shape[2]
    xbar = s[:, 0].copy()
    for j in range(1, N):  # np.add.reduce(sim[:-1], 0): sequential row sum
        xbar = xbar + s[:, j]
    xbar = xbar / N
    last = s[:, -1]
    xr = (1 + RHO) * xbar - RHO * last
    xe = (1 + RHO * CHI) * xbar - RHO * CHI * last
    xc = (1 + PSI * RHO) * xbar - PSI * RHO * last
    xcc = (1 - PSI) * xbar + PSI * last
    return xr, xe, xc, xcc


def _decide(f, fxr, fxe, fxc, fxcc):
    """scipy's decision tree (_optimize.py:846-896), vectorised over simplices: which candidate replaces the worst vertex
    (0 expansion, 1 reflection, 2 outside contraction, 3 inside contraction, -1 none = shrink) and scipy's evaluation count"""
    better_than_best = fxr < f[:, 0]
    take_e = better_than_best & (fxe < fxr)
    take_r = (better_than_best & ~(fxe < fxr)) | (~better_than_best & (fxr < f[:, -2]))
    contract = ~better_than_best & ~(fxr < f[:, -2])
    outside = contract & (fxr < f[:, -1])
    inside = contract & ~(fxr < f[:, -1])
    take_c = outside & (fxc <= fxr)
    take_cc = inside & (fxcc < f[:, -1])
    which = np.where(take_e, 0, np.where(take_r, 1, np.where(take_c, 2, np.where(take_cc, 3, -1))))
    nev = 1 + (better_than_best | contract).astype(np.int64)
    return which, nev


def nelder_mead_batch(fun, x0, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, speculative=True, owners=None, lookahead="auto"):
    """Minimise S independent objectives in lock step.  x0: [S, N].  Returns a dict of arrays:
    x [S, N], fun [S], nit [S], nfev [S] (scipy's count), status [S] (0 converged, 1 maxfev, 2 maxiter),
    success [S], evaluations (points actually sent to `fun`), launches (calls of `fun`).

    lookahead: with the candidates of the current step also evaluate, speculatively, the candidates of the NEXT step for
    every way the current one can end (which candidate is accepted x where it lands in the sorted simplex: 3N + 3
    scenarios of 4 points), so that one call of `fun` advances a simplex by two iterations.  The decisions and scipy's
    counts are unchanged; it trades (12N + 12) extra points per call for half the calls, which pays while a call's cost
    does not depend on its size (the device evaluates a few thousand points as fast as one).  "auto" = while the
    call stays below LOOKAHEAD_MAX_POINTS points."""
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

    if maxfev < N + 1:  # scipy stops evaluating the initial simplex when the budget is spent; the rest stays at +inf
        n0 = int(maxfev)
        fsim = np.full((S, N + 1), np.inf)
        if n0 > 0:
            fsim[:, :n0] = call(sim[:, :n0].reshape(-1, N), np.repeat(np.arange(S), n0)).reshape(S, n0)
        fcalls = np.full(S, n0, dtype=np.int64)
    else:
        fsim = call(sim.reshape(-1, N), np.repeat(np.arange(S), N + 1)).reshape(S, N + 1)
        fcalls = np.full(S, N + 1, dtype=np.int64)
    sim, fsim = _sort(sim, fsim)
    iterations = np.ones(S, dtype=np.int64)
    status = np.full(S, -1, dtype=np.int64)
    active = np.ones(S, dtype=bool)
    # scenarios of the look-ahead: (accepted candidate, rank of the new vertex in the sorted simplex); see _decide for
    # the conditions that make other combinations impossible
    scen = [(0, 0)] + [(1, k) for k in range(N)] + [(2, k) for k in range(N + 1)] + [(3, k) for k in range(N + 1)]
    n_scen = len(scen)
    scen_base = np.array([0, 1, 1 + N, 2 + 2 * N], dtype=np.int64)  # first scenario of each accepted candidate

    def retire():
        """termination tests at the top of scipy's loop; returns the indices still active"""
        budget = (fcalls < maxfev) & (iterations < maxiter)
        with np.errstate(invalid="ignore"):
            conv = (np.max(np.abs(sim[:, 1:] - sim[:, :1]).reshape(S, -1), axis=1) <= xatol) & \
                   (np.max(np.abs(fsim[:, :1] - fsim[:, 1:]), axis=1) <= fatol)
        done_now = active & (~budget | conv)
        status[done_now & budget & conv] = 0
        status[done_now & ~budget & (fcalls >= maxfev)] = 1
        status[done_now & ~budget & (fcalls < maxfev)] = 2
        active[done_now] = False
        return np.nonzero(active)[0]

    def apply(idx, s, f, which, nev, cand, fcand):
        """replace the worst vertex by the accepted candidate, or shrink (one more call); sort; write back"""
        nonlocal sim, fsim
        A = len(idx)
        rows = np.arange(A)
        acc = which >= 0
        w = np.where(acc, which, 0)
        new_x = cand[w, rows]
        new_f = fcand[w, rows]
        s[acc, -1] = new_x[acc]
        f[acc, -1] = new_f[acc]
        fcalls[idx] += nev
        shrink = ~acc
        if shrink.any():
            sh = np.nonzero(shrink)[0]
            s[sh, 1:] = s[sh, :1] + SIGMA * (s[sh, 1:] - s[sh, :1])
            fs = call(s[sh, 1:].reshape(-1, N), np.repeat(idx[sh], N)).reshape(len(sh), N)
            f[sh, 1:] = fs
            fcalls[idx[sh]] += N
        s, f = _sort(s, f)
        sim[idx], fsim[idx] = s, f
        iterations[idx] += 1

    class _BudgetSpent(Exception):
        pass

    def finish_serially(i):
        """scipy enforces maxfev INSIDE an iteration (its objective wrapper raises once the budget is spent,
        _optimize.py:549-559, and the iteration is abandoned where it stands).  A simplex that could run out of budget
        within the coming step is therefore taken through scipy's own loop (_optimize.py:867-934), one evaluation at a
        time, to its end."""
        def f(x):
            if fcalls[i] >= maxfev:
                raise _BudgetSpent
            fcalls[i] += 1
            return call(x.reshape(1, N), np.array([i]))[0]
        s, fs = sim[i].copy(), fsim[i].copy()
        while fcalls[i] < maxfev and iterations[i] < maxiter:
            try:
                with np.errstate(invalid="ignore"):
                    if np.max(np.abs(s[1:] - s[0])) <= xatol and np.max(np.abs(fs[0] - fs[1:])) <= fatol:
                        break
                xr, xe, xc, xcc = (v[0] for v in _candidates(s[None]))
                fxr = f(xr)
                doshrink = False
                if fxr < fs[0]:
                    fxe = f(xe)
                    if fxe < fxr:
                        s[-1], fs[-1] = xe, fxe
                    else:
                        s[-1], fs[-1] = xr, fxr
                elif fxr < fs[-2]:
                    s[-1], fs[-1] = xr, fxr
                elif fxr < fs[-1]:
                    fxc = f(xc)
                    if fxc <= fxr:
                        s[-1], fs[-1] = xc, fxc
                    else:
                        doshrink = True
                else:
                    fxcc = f(xcc)
                    if fxcc < fs[-1]:
                        s[-1], fs[-1] = xcc, fxcc
                    else:
                        doshrink = True
                if doshrink:
                    for j in range(1, N + 1):
                        s[j] = s[0] + SIGMA * (s[j] - s[0])
                        fs[j] = f(s[j])
                iterations[i] += 1
            except _BudgetSpent:
                pass
            ind = np.argsort(fs, kind="stable")
            s, fs = s[ind], fs[ind]
        sim[i], fsim[i] = s, fs
        status[i] = 1 if fcalls[i] >= maxfev else (2 if iterations[i] >= maxiter else 0)
        active[i] = False

    while True:
        idx = retire()
        if maxfev != np.inf and idx.size:
            for i in idx[fcalls[idx] + 2 * (2 + N) > maxfev]:  # a call may take two iterations (look-ahead)
                finish_serially(int(i))
            idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        A = idx.size
        s, f = sim[idx].copy(), fsim[idx].copy()
        xr, xe, xc, xcc = _candidates(s)
        cand = np.stack([xe, xr, xc, xcc])  # order of _decide's codes
        look = speculative and (lookahead is True or (lookahead == "auto" and A * 4 * (1 + n_scen) <= LOOKAHEAD_MAX_POINTS))
        if not speculative:
            fxr = call(xr, idx)
            want_e = fxr < f[:, 0]
            want_c = ~want_e & ~(fxr < f[:, -2]) & (fxr < f[:, -1])
            want_cc = ~want_e & ~(fxr < f[:, -2]) & ~(fxr < f[:, -1])
            second = np.where(want_e[:, None], xe, np.where(want_c[:, None], xc, xcc))
            need = want_e | want_c | want_cc
            f2 = np.full(A, np.inf)
            if need.any():
                f2[need] = call(second[need], idx[need])
            fcand = np.stack([f2, fxr, f2, f2])
            which, nev = _decide(f, fxr, f2, f2, f2)
            apply(idx, s, f, which, nev, cand, fcand)
            continue
        if not look:
            fcand = call(np.concatenate([xe, xr, xc, xcc]), np.tile(idx, 4)).reshape(4, A)
            which, nev = _decide(f, fcand[1], fcand[0], fcand[2], fcand[3])
            apply(idx, s, f, which, nev, cand, fcand)
            continue
        # ---- two iterations per call
        nxt = np.empty((n_scen, 4, A, N))  # candidates of the next step per scenario, in _decide's order
        nsim = np.empty((n_scen, A, N + 1, N))
        for q, (o, k) in enumerate(scen):
            t = np.empty((A, N + 1, N))
            t[:, :k] = s[:, :k]
            t[:, k] = cand[o]
            t[:, k + 1:] = s[:, k:N]
            nsim[q] = t
            r2, e2, c2, cc2 = _candidates(t)
            nxt[q, 0], nxt[q, 1], nxt[q, 2], nxt[q, 3] = e2, r2, c2, cc2
        X = np.concatenate([cand.reshape(4 * A, N), nxt.reshape(n_scen * 4 * A, N)])
        fall = call(X, np.tile(idx, 4 * (1 + n_scen)))
        fcand = fall[:4 * A].reshape(4, A)
        fnxt = fall[4 * A:].reshape(n_scen, 4, A)
        which, nev = _decide(f, fcand[1], fcand[0], fcand[2], fcand[3])
        apply(idx, s, f, which, nev, cand, fcand)
        # second iteration for the simplices that are still active and whose first step was not a shrink
        keep = which >= 0
        retire()
        sel = np.nonzero(active[idx] & keep)[0]
        if sel.size == 0:
            continue
        idx2 = idx[sel]
        s2, f2 = sim[idx2].copy(), fsim[idx2].copy()
        # the scenario that came true: the accepted candidate and the rank it got in the stable sort
        o = which[sel]
        fnew = fcand[o, sel]
        k = (f[sel][:, :N] <= fnew[:, None]).sum(axis=1)  # f still holds the pre-sort values with the new one last
        q = scen_base[o] + k  # index of (o, k) in scen
        assert np.array_equal(nsim[q, sel], s2), "look-ahead scenario does not match the simplex"
        cand2 = nxt[q, :, sel].transpose(1, 0, 2)  # [4, A2, N]
        fcand2 = fnxt[q, :, sel].T                 # [4, A2]
        which2, nev2 = _decide(f2, fcand2[1], fcand2[0], fcand2[2], fcand2[3])
        apply(idx2, s2, f2, which2, nev2, cand2, fcand2)
    return {"x": sim[:, 0].copy(), "fun": fsim.min(axis=1), "nit": iterations, "nfev": fcalls, "status": status,
            "success": status == 0, "sim": sim, "fsim": fsim, "evaluations": evaluations, "launches": launches}


def basinhopping_batch(fun, x0, niter=100, T=1.0, stepsize=0.5, interval=50, target_accept_rate=0.5, stepwise_factor=0.9,
                       seeds=None, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, speculative=True, local_solver=None):
    """W basin-hopping walkers in lock step (scipy/optimize/_basinhopping.py), local search = nelder_mead_batch.
    x0: [W, N]; seeds: one seed (or numpy Generator) per walker.  Returns dict: x [W, N], fun [W], nfev [W],
    nit, accepted [W], minimization_failures [W], evaluations, launches.
    local_solver(xs[W, N]) -> the dict of nelder_mead_batch: replaces the host-driven local search (e.g. by the on-device
    Nelder-Mead, Engine.nelder_mead); `fun` is then not called."""
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
        if local_solver is not None:
            r = local_solver(xs)
        else:
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
            with np.errstate(invalid="ignore"):  # inf - inf = nan: min(0, nan) is 0 in Python, as in scipy
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
