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
