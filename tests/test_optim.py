"""Batched optimisers (misti_b200/optim.py) against scipy's serial drivers on analytic objectives (CPU):
identical decisions => identical x, f, iteration and evaluation counts for every simplex of a batch."""
import numpy as np
import pytest
from scipy import optimize

from misti_b200.optim import basinhopping_batch, initial_simplex, nelder_mead_batch


def rosen_like(X, shift):
    X = np.atleast_2d(X)
    a = X[:, 0] - shift
    f = 100.0 * (X[:, 1] - a * a) ** 2 + (1 - a) ** 2
    f = np.where((X < 0).any(axis=1), np.inf, f)  # negative parameters -> +inf, like the objective (MigrationInference.py:569-572)
    return f


def bumpy(X, shift):
    X = np.atleast_2d(X)
    return np.cos(14.5 * (X[:, 0] - shift) - 0.3) + ((X[:, 0] - shift) + 0.2) * (X[:, 0] - shift) + 0.1 * (X ** 2).sum(axis=1)


@pytest.mark.parametrize("speculative", [True, False])
def test_nelder_mead_batch_equals_scipy(speculative):
    rng = np.random.default_rng(0)
    S = 23
    x0 = rng.uniform(0.0, 3.0, (S, 2))
    x0[3, 1] = 0.0  # zero coordinate -> 0.00025 step
    shifts = rng.uniform(0.0, 1.0, S)
    res = nelder_mead_batch(lambda X, who: rosen_like(X, shifts[who]), x0, xatol=1e-4, fatol=1e-4, maxiter=1000,
                            speculative=speculative)
    for s in range(S):
        ref = optimize.minimize(lambda x: float(rosen_like(x, shifts[s])[0]), x0[s], method="Nelder-Mead",
                                options={"xatol": 1e-4, "fatol": 1e-4, "maxiter": 1000})
        assert np.array_equal(res["x"][s], ref.x), s
        assert res["fun"][s] == ref.fun
        assert res["nit"][s] == ref.nit and res["nfev"][s] == ref.nfev
        assert bool(res["success"][s]) == ref.success
    assert res["launches"] < 200 * (1 if speculative else 2)


def test_iteration_budget_and_initial_simplex():
    x0 = np.array([[1.3, 0.0, 2.0]])
    sim = initial_simplex(x0)[0]
    assert np.array_equal(sim[1], [1.3 * 1.05, 0.0, 2.0]) and sim[2][1] == 0.00025 and sim[3][2] == 2.0 * 1.05
    f = lambda X, who: ((np.atleast_2d(X) - 0.7) ** 2).sum(axis=1)  # noqa: E731
    res = nelder_mead_batch(f, x0, maxiter=7)
    ref = optimize.minimize(lambda x: float(f(x, None)[0]), x0[0], method="Nelder-Mead", options={"maxiter": 7})
    assert res["status"][0] == 2 and not ref.success
    assert np.array_equal(res["x"][0], ref.x) and res["nit"][0] == ref.nit and res["nfev"][0] == ref.nfev
    # scipy defaults (both budgets = 200 N), as basin-hopping's inner minimiser uses them
    res = nelder_mead_batch(f, x0)
    ref = optimize.minimize(lambda x: float(f(x, None)[0]), x0[0], method="Nelder-Mead")
    assert np.array_equal(res["x"][0], ref.x) and res["nfev"][0] == ref.nfev


def test_all_infinite_objective_terminates():
    res = nelder_mead_batch(lambda X, who: np.full(len(X), np.nan), np.array([[1.0, 2.0]]), maxiter=50)
    assert res["fun"][0] == np.inf and res["status"][0] == 2


def test_basinhopping_walkers_equal_scipy():
    W = 3
    x0 = np.array([[1.0], [0.2], [2.5]])
    shifts = np.array([0.0, 0.3, 0.7])
    res = basinhopping_batch(lambda X, who: bumpy(X, shifts[who]), x0, niter=25, T=0.5, stepsize=0.5, interval=10,
                             seeds=[2024, 2025, 2026])
    for w in range(W):
        ref = optimize.basinhopping(lambda x: float(bumpy(x, shifts[w])[0]), x0[w], niter=25, T=0.5, stepsize=0.5, interval=10,
                                    minimizer_kwargs=dict(method="Nelder-Mead"), rng=np.random.default_rng(2024 + w))
        assert np.array_equal(res["x"][w], ref.x), w
        assert res["fun"][w] == ref.fun
        assert res["nfev"][w] == ref.nfev
        assert res["minimization_failures"][w] == ref.minimization_failures


@pytest.mark.parametrize("N", [1, 2, 3])
def test_lookahead_takes_the_same_decisions_in_half_the_calls(N):
    """two Nelder-Mead iterations per objective call (speculative candidates of the next step for every possible outcome
    of the current one): same x, f, nit, nfev as scipy, about half the calls"""
    rng = np.random.default_rng(N)
    S = 17
    x0 = rng.uniform(0.2, 3.0, (S, N))
    shifts = rng.uniform(0.0, 1.0, S)

    def obj(X, who):
        X = np.atleast_2d(X)
        a = X - shifts[np.asarray(who)][:, None]
        f = (a ** 2).sum(axis=1) + 0.3 * np.cos(3.0 * a).sum(axis=1) + (0.5 * (a[:, :1] * a[:, -1:]).sum(axis=1) if N > 1 else 0.0)
        return np.where((X < 0).any(axis=1), np.inf, f)
    res = {la: nelder_mead_batch(obj, x0, xatol=1e-6, fatol=1e-6, maxiter=400, lookahead=la) for la in (False, True)}
    for k in ("x", "fun", "nit", "nfev", "status"):
        assert np.array_equal(res[False][k], res[True][k]), k
    assert res[True]["launches"] < 0.62 * res[False]["launches"]
    for s in range(S):
        ref = optimize.minimize(lambda x: float(obj(x, [s])[0]), x0[s], method="Nelder-Mead",
                                options={"xatol": 1e-6, "fatol": 1e-6, "maxiter": 400})
        assert np.array_equal(res[True]["x"][s], ref.x) and res[True]["nfev"][s] == ref.nfev and res[True]["nit"][s] == ref.nit


def _device_logic_nelder_mead(lib, fun, x0, xatol=1e-4, fatol=1e-4, maxiter=-1, maxfev=-1, lookahead=0):
    """One simplex through the step logic of the ON-DEVICE Nelder-Mead (misti_b200/csrc/misti_optim.cuh, compiled for the
    host by tests/hostsim): propose -> evaluate here -> apply, until nothing is proposed."""
    import ctypes
    dp, lp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)
    N = len(x0)
    sim, fsim = np.zeros((N + 1, N)), np.zeros(N + 1)
    sim[0] = x0
    iters, fcalls = ctypes.c_longlong(0), ctypes.c_longlong(0)
    status, phase = ctypes.c_int(-1), ctypes.c_int(0)
    pts = np.zeros((lib.hs_nm_slots(N, lookahead), N))
    cfg = (N, ctypes.c_double(xatol), ctypes.c_double(fatol), ctypes.c_longlong(maxiter), ctypes.c_longlong(maxfev), lookahead)
    lib.hs_nm_propose.restype = ctypes.c_int
    rounds = 0
    while True:
        n = lib.hs_nm_propose(*cfg, sim.ctypes.data_as(dp), fsim.ctypes.data_as(dp), ctypes.byref(iters), ctypes.byref(fcalls),
                              ctypes.byref(status), ctypes.byref(phase), pts.ctypes.data_as(dp))
        if n == 0:
            break
        fv = np.full(len(pts), np.nan)
        fv[:n] = [fun(pts[j].copy()) for j in range(n)]
        lib.hs_nm_apply(*cfg, sim.ctypes.data_as(dp), fsim.ctypes.data_as(dp), ctypes.byref(iters), ctypes.byref(fcalls),
                        ctypes.byref(status), ctypes.byref(phase), pts.ctypes.data_as(dp), fv.ctypes.data_as(dp))
        rounds += 1
        assert rounds < 100000
    return dict(x=sim[0].copy(), fun=fsim[0], nit=iters.value, nfev=fcalls.value, status=status.value, rounds=rounds)


@pytest.mark.parametrize("N,lookahead", [(1, 0), (2, 0), (3, 0), (5, 0), (1, 1), (2, 1), (3, 1), (4, 1)])
def test_device_step_logic_equals_scipy(hostsim, N, lookahead):
    """the decisions, iterates and counts of the on-device Nelder-Mead are scipy's, also through +inf / NaN objective
    values, shrinks, and the iteration / evaluation budgets"""
    rng = np.random.default_rng(10 + N)

    def make(shift, kind):
        def f(x):
            x = np.asarray(x, dtype=np.float64)
            if kind == 0:  # smooth valley with an infeasible region
                if (x < 0).any():
                    return np.inf
                a = x[0] - shift
                b = x[1 % len(x)]
                return float(100.0 * (b - a * a) ** 2 + (1 - a) ** 2 + 0.1 * (x ** 2).sum())
            if kind == 1:  # bumpy: many contractions and shrinks
                return float(np.cos(14.5 * (x[0] - shift) - 0.3) + ((x[0] - shift) + 0.2) * (x[0] - shift) + 0.1 * (x ** 2).sum()
                             + np.abs(np.sin(37.0 * x).sum()))
            return float("nan") if x[0] > 2.5 else float(((x - shift) ** 2).sum())  # NaN counts as +inf
        return f
    shrinks = 0
    for trial in range(24):
        x0 = rng.uniform(0.0, 3.0, N)
        if trial % 5 == 0:
            x0[-1] = 0.0  # zero coordinate -> 0.00025 step
        fun = make(rng.uniform(0, 1), trial % 3)
        budget = {} if trial % 4 else {"maxiter": 15 + trial}
        if trial % 6 == 1:
            budget = {"maxfev": 20 + trial}
        opts = {"xatol": 1e-4, "fatol": 1e-4}
        opts.update(budget)
        with np.errstate(invalid="ignore"):
            clean = (lambda x: np.inf if np.isnan(fun(x)) else fun(x))
            ref = optimize.minimize(clean, x0, method="Nelder-Mead", options=opts)
        # scipy's defaults for the missing budget (_optimize.py:751-768), as Engine.nelder_mead applies them
        mi, mf = budget.get("maxiter"), budget.get("maxfev")
        if mi is None and mf is None:
            mi, mf = N * 200, N * 200
        elif mi is None:
            mi = -1
        elif mf is None:
            mf = -1
        got = _device_logic_nelder_mead(hostsim, fun, x0, maxiter=mi, maxfev=mf, lookahead=lookahead)
        assert np.array_equal(got["x"], ref.x), (N, trial)
        assert got["rounds"] <= (ref.nit if not lookahead else ref.nit // 2 + 1) + (got["nfev"] - ref.nit) + 2
        assert got["fun"] == ref.fun and got["nit"] == ref.nit and got["nfev"] == ref.nfev, (N, trial, got, ref.nit, ref.nfev)
        assert got["status"] == ref.status, (N, trial)
        shrinks += got["nfev"] > 2 * got["nit"] + N + 1
    assert N == 1 or shrinks > 0


def test_host_driver_follows_scipy_through_the_evaluation_budget():
    """maxfev is enforced inside an iteration by scipy (the iteration is abandoned where it stands); the lock-step
    driver hands a simplex that is about to run out to scipy's own serial loop, so x, counts and status stay scipy's
    for every budget, with and without look-ahead"""
    def f1(x):
        x = np.asarray(x)
        return float(np.cos(14.5 * x[0] - 0.3) + (x[0] + 0.2) * x[0] + 0.1 * (x ** 2).sum() + abs(np.sin(37 * x).sum()))
    rng = np.random.default_rng(3)
    for N in (1, 2, 3):
        for mf in list(range(1, 26)) + [N * 200]:
            for look, spec in ((False, True), ("auto", True), (False, False)):
                x0 = rng.uniform(0, 3, (3, N))
                r = nelder_mead_batch(lambda X, who: np.array([f1(x) for x in X]), x0, maxfev=mf, lookahead=look, speculative=spec)
                for s in range(3):
                    ref = optimize.minimize(f1, x0[s], method="Nelder-Mead", options={"maxfev": mf, "xatol": 1e-4, "fatol": 1e-4})
                    assert np.array_equal(r["x"][s], ref.x), (N, mf, look, spec)
                    assert (r["nfev"][s], r["nit"][s], r["status"][s]) == (ref.nfev, ref.nit, ref.status), (N, mf, look, spec)


def _pcg_state(seed):
    st = np.random.default_rng(seed).bit_generator.state["state"]
    mask = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & mask, st["inc"] >> 64, st["inc"] & mask], dtype=np.uint64)


def test_device_generator_continues_numpys_stream(hostsim):
    """the walkers' random numbers (misti_optim.cuh: Pcg64) are numpy's Generator(PCG64) stream, draw for draw:
    uniform(-s, s, N) displacements and uniform() Metropolis numbers in any interleaving"""
    import ctypes
    up, dp = ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_double)
    for seed in (0, 1, 2024, 123456789):
        ref = np.random.default_rng(seed)
        state = _pcg_state(seed)
        for low, high, n in ((-0.5, 0.5, 3), (0.0, 1.0, 1), (-0.45, 0.45, 5), (0.0, 1.0, 1), (-1.7, 1.7, 64)):
            out = np.zeros(n)
            hostsim.hs_pcg64_uniform(state.ctypes.data_as(up), ctypes.c_double(low), ctypes.c_double(high), n, out.ctypes.data_as(dp))
            want = ref.uniform(low, high, n) if n > 1 else np.array([ref.uniform(low, high)])
            assert np.array_equal(out, want), (seed, low, high)


def _device_logic_basinhopping(lib, fun, x0, seed, niter, T=0.5, stepsize=0.5, interval=50, target=0.5, factor=0.9):
    """one walker through the device's walker logic (bh_advance around the Nelder-Mead step logic), objective on the host"""
    import ctypes
    dp, lp, ip, up = (ctypes.POINTER(t) for t in (ctypes.c_double, ctypes.c_longlong, ctypes.c_int, ctypes.c_ulonglong))
    N = len(x0)
    sim, fsim = np.zeros((N + 1, N)), np.zeros(N + 1)
    sim[0] = x0
    iters, fcalls = ctypes.c_longlong(0), ctypes.c_longlong(0)
    status, phase = ctypes.c_int(-1), ctypes.c_int(0)
    pts = np.zeros((lib.hs_nm_slots(N, 0), N))
    cfg = (N, ctypes.c_double(1e-4), ctypes.c_double(1e-4), ctypes.c_longlong(200 * N), ctypes.c_longlong(200 * N), 0)
    x, best_x, scal = np.zeros(N), np.zeros(N), np.array([0.0, 0.0, stepsize])
    flags, cnt, rng = np.zeros(3, dtype=np.int32), np.zeros(5, dtype=np.int64), _pcg_state(seed)
    lib.hs_nm_propose.restype = ctypes.c_int
    while True:
        n = lib.hs_nm_propose(*cfg, sim.ctypes.data_as(dp), fsim.ctypes.data_as(dp), ctypes.byref(iters), ctypes.byref(fcalls),
                              ctypes.byref(status), ctypes.byref(phase), pts.ctypes.data_as(dp))
        if n == 0:  # the local search has ended: the walker decides and, unless it has done its hops, starts the next one
            go = lib.hs_bh_advance(N, niter, interval, ctypes.c_double(T), ctypes.c_double(target), ctypes.c_double(factor),
                                   x.ctypes.data_as(dp), best_x.ctypes.data_as(dp), scal.ctypes.data_as(dp), flags.ctypes.data_as(ip),
                                   cnt.ctypes.data_as(lp), rng.ctypes.data_as(up), sim.ctypes.data_as(dp), fsim.ctypes.data_as(dp),
                                   ctypes.byref(iters), ctypes.byref(fcalls), ctypes.byref(status), ctypes.byref(phase))
            if not go:
                break
            continue
        fv = np.full(len(pts), np.nan)
        fv[:n] = [fun(pts[j].copy()) for j in range(n)]
        lib.hs_nm_apply(*cfg, sim.ctypes.data_as(dp), fsim.ctypes.data_as(dp), ctypes.byref(iters), ctypes.byref(fcalls),
                        ctypes.byref(status), ctypes.byref(phase), pts.ctypes.data_as(dp), fv.ctypes.data_as(dp))
    return dict(x=best_x.copy(), fun=scal[1], nfev=int(cnt[0]), failures=int(cnt[1]), accepted=int(cnt[3]), step=scal[2])


@pytest.mark.parametrize("N", [1, 2, 3])
def test_device_walker_logic_equals_scipy_basinhopping(hostsim, N):
    """a single walker of the on-device basin-hopping reproduces scipy.optimize.basinhopping(..., rng=seed) as
    MigrationInference.Solve(globalOpt=True) calls it (MigrationInference.py:724): same minimum, same evaluation count,
    same number of failed local searches -- through adaptive step sizes (interval 7) and infinite objective values"""
    from scipy import optimize
    rng = np.random.default_rng(40 + N)
    for trial in range(3):
        shift = rng.uniform(-1, 1, N)

        def f(x, shift=shift):
            if x[0] < -1.2:  # a region where the model does not evaluate (negative rate): +inf, as the reference returns
                return np.inf
            return float(bumpy(np.asarray(x).reshape(1, -1), shift)[0])
        x0 = rng.uniform(-1, 1, N)
        seed = 1000 * N + trial
        got = _device_logic_basinhopping(hostsim, f, x0, seed, niter=12, T=0.5, stepsize=0.5, interval=7)
        ref = optimize.basinhopping(f, x0, niter=12, T=0.5, stepsize=0.5, interval=7, rng=seed,
                                    minimizer_kwargs={"method": "Nelder-Mead"})
        assert np.array_equal(got["x"], ref.x) and got["fun"] == ref.fun, (N, trial)
        assert got["nfev"] == ref.nfev and got["failures"] == ref.minimization_failures, (N, trial)
