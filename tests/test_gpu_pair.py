"""GPU parity of the pair-of-lanes JSFS kernel (misti_jsfs_pair_kernel, csrc/misti_pair.cuh): the kernel large batches
run (B > 6 144), forced here for single items as well (MISTI_JSFS_PAIR = 1) so that the golden vectors of the unmodified
reference can be held against it, and compared on large mixed batches with the 16-lane kernel it replaces.
Tolerance against the reference: 1e-9 relative on every JSFS entry and on llh (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from _cases import RUNAWAY, bands_pulses, end_to_end_gated, flags_of, grid_of, relerr, sfs_of

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _engine(**env):
    import misti_b200
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return misti_b200.Engine(0)  # the tuning knobs are read when the context is created
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def pair_engine():
    eng = _engine(MISTI_JSFS_PAIR=1, MISTI_DEFER_POST=2)
    yield eng
    eng.close()


def _register(engine, ds, case):
    times, lam, st, sd = grid_of(ds, case)
    bands, pulses = bands_pulses(case)
    engine.clear_models()
    gid = engine.add_grid(times, lam)
    mid = engine.add_model(gid, st, sd, bands, pulses)
    engine.set_data([sfs_of(ds, case)], case["flags"]["unfolded"])
    return mid, len(lam)


def test_pair_kernel_golden_jsfs_stage(pair_engine, golden_datasets, golden_cases):
    """JSFS + likelihood given the reference's corrected rates, every mode: bands, pulses, ancient sample, folded / unfolded."""
    n = 0
    for case in golden_cases:
        exp = case["expect"]
        if not exp["ok"] or case["name"] in RUNAWAY:
            continue
        mid, numT = _register(pair_engine, golden_datasets, case)
        P = len(case["params"])
        inj = np.zeros((1, pair_engine.numT_max, 2))
        inj[0, :numT] = np.array(exp["lc"])
        out = pair_engine.evaluate(np.array([case["params"]]).reshape(1, P), model=mid, flags=flags_of(case), lc_inject=inj,
                                   want=("jafs", "status", "terms"))
        assert out["status"][0] == 0, case["name"]
        assert relerr(out["jafs"][0], exp["JAFS"]) < TOL, case["name"]
        assert relerr(out["llh"][0, 0], exp["llh"]) < TOL, case["name"]
        n += 1
    assert n >= 40


def test_pair_kernel_golden_end_to_end(pair_engine, golden_datasets, golden_cases):
    """correction chain + post-split kernel + pair kernel vs the reference's outputs."""
    checked = 0
    for case in golden_cases:
        mid, numT = _register(pair_engine, golden_datasets, case)
        P = len(case["params"])
        out = pair_engine.evaluate(np.array([case["params"]]).reshape(1, P), model=mid, flags=flags_of(case), want=("jafs", "status"))
        exp = case["expect"]
        if not exp["ok"]:
            assert out["status"][0] in (1, 2), case["name"]
            assert out["llh"][0, 0] == -np.inf, case["name"]
            continue
        if not end_to_end_gated(case):
            continue
        assert out["status"][0] == 0, (case["name"], out["status"])
        assert relerr(out["jafs"][0], exp["JAFS"]) < TOL, case["name"]
        assert relerr(out["llh"][0, 0], exp["llh"]) < TOL, case["name"]
        checked += 1
    assert checked >= 35


def test_large_batches_take_the_pair_kernel_and_agree_with_the_16_lane_kernel(golden_datasets):
    """B > 6 144 with the default knobs against MISTI_JSFS_PAIR = 0: mixed models in one batch (warps whose items differ
    in model, items with stiff segments and a model without a split inside the grid go to the 16-lane kernel through
    the redo list), negative parameters, few and many data rows.  Same status and term counts; numbers to rounding."""
    ds = golden_datasets["synthetic"]
    numT = len(ds["lambdas"])
    rng = np.random.default_rng(5)
    n = 24000
    p = np.column_stack([rng.uniform(0, 3, n), rng.uniform(0, 3, n), rng.uniform(0, 0.5, n)])
    p[::101, 0] = -0.5
    res = []
    for knob in (None, 0):
        eng = _engine(**({} if knob is None else {"MISTI_JSFS_PAIR": knob}))
        gid = eng.add_grid(ds["times"], ds["lambdas"])
        ms = [eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)]),
              eng.add_model(gid, 38, 0, bands=[(0, 4, 38, 3.0, 0)]),  # stiff for large rates
              eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)]),
              eng.add_model(gid, 44, 12, bands=[(1, 5, 20, 0.8, 0)], pulses=[(1, 12, 0.1, -1)]),
              eng.add_model(gid, 30, 30, bands=[(0, 3, 9, 0.5, 0)]),
              eng.add_model(gid, numT, 0, bands=[(0, 100, numT, 0.5, 0), (1, 100, numT, 0.7, -1)])]  # infinite last interval
        # runs of 1 500 items per model (homogeneous warps) followed by items whose models alternate
        mids = np.array(ms, dtype=np.int32)[np.concatenate([np.repeat(np.arange(6), 1500) % 6, np.arange(n - 9000) % 6])]
        outs = {}
        for R in (1, 5, 40, 70):  # staged stores (<= 32 rows), direct stores, the scoring kernel (>= 64)
            rows = [ds["sfs"]] + [list(r) for r in ds.get("bs_rows", [])]
            while len(rows) < R:
                rows.append([v * (1.0 + 0.01 * len(rows)) for v in ds["sfs"]])
            eng.set_data(rows[:R], True)
            outs[R] = eng.evaluate(p, model_ids=mids, flags=15, want=("jafs", "status", "terms"))
        eng.set_data([ds["sfs"]], True)
        outs["one_model"] = eng.evaluate(p[:, :1].copy(), model=ms[0], flags=15, want=("jafs", "status", "terms"))
        # a split-time grid interleaved item by item: segment lists of the same type pattern, lock step across models
        grid = np.array([eng.add_model(gid, st, 0, bands=[(1, 5, 12, 0.8, 0)]) for st in range(36, 45)], dtype=np.int32)
        outs["split_grid"] = eng.evaluate(p, model_ids=grid[np.arange(n) % 9], flags=15, want=("jafs", "status", "terms"))
        res.append(outs)
        eng.close()
    for key in res[0]:
        a, b = res[0][key], res[1][key]
        assert np.array_equal(a["status"], b["status"]), key
        assert np.array_equal(a["terms"], b["terms"]), key
        assert (a["status"] == 0).sum() > 0.5 * n and (a["status"] != 0).sum() > 100, key
        for k in ("llh", "jafs"):
            assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])) and np.array_equal(np.isinf(a[k]), np.isinf(b[k])), (key, k)
            m = np.isfinite(a[k])
            assert float(np.max(np.abs(a[k][m] - b[k][m]) / np.abs(b[k][m]))) < 1e-10, (key, k)


def test_one_block_per_sm_launches_of_the_correction_kernel_change_nothing(golden_datasets):
    """Batches that fill the machine launch the correction kernel as one block per SM with a barrier at every interval of
    the chain (csrc/misti_kernels.cu: correct_big_blocks / correct_align).  The arithmetic per thread is the same, so the
    results are those of the two-warp blocks bit for bit -- also when threads leave the chain early (negative parameters,
    corrections that fail in the reference's default mode), when models of different length share a block, and when the
    batch is not a multiple of the block."""
    ds = golden_datasets["synthetic"]
    numT = len(ds["lambdas"])
    rng = np.random.default_rng(21)
    n = 60001
    p = rng.uniform(0, 5, (n, 1))
    p[::53, 0] = -1.0
    pick = rng.integers(0, 5, n)
    res = []
    for env in ({}, {"MISTI_CORRECT_BIG_BLOCKS": 0}, {"MISTI_CORRECT_ALIGN": 0}):
        eng = _engine(**env)
        gid = eng.add_grid(ds["times"], ds["lambdas"])
        ms = np.array([eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)]), eng.add_model(gid, 20, 0, bands=[(1, 5, 12, 0.8, 0)]),
                       eng.add_model(gid, 90, 0, bands=[(0, 4, 38, 3.0, 0)]), eng.add_model(gid, numT, 0, bands=[(0, 100, numT, 0.5, 0)]),
                       eng.add_model(gid, 36, 0)], dtype=np.int32)
        eng.set_data([ds["sfs"]], True)
        res.append([eng.evaluate(p[:33000], model_ids=ms[pick[:33000]], flags=15, want=("jafs", "status", "nfev", "terms")),  # blocks of two warps
                    eng.evaluate(p, model=int(ms[0]), flags=15, want=("jafs", "status", "nfev", "terms")),
                    eng.evaluate(p, model_ids=ms[pick], flags=15, want=("jafs", "status", "nfev", "terms")),
                    eng.evaluate(p, model_ids=ms[pick], flags=13, want=("jafs", "status", "nfev", "terms"))])
        eng.close()
    for other in res[1:]:
        for a, b in zip(res[0], other):
            for k in a:
                assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert (res[0][3]["status"] != 0).sum() > 1000 and (res[0][2]["status"] == 0).sum() > 1000


def test_large_batch_path_against_the_oracle(golden_datasets):
    """The kernels a machine-filling batch runs by default -- correction kernel as one block per SM with interval barriers,
    four-lane post-split kernel, pair-of-lanes JSFS kernel -- held DIRECTLY against the CPU oracle on items taken from a
    60 000-item batch of BASELINE config 2 (`-uf -mi 2 5 12 0.8 1 --cpfit`) and of config 3 (two bands + pulse)."""
    import misti_b200
    from oracle.misti_oracle import OracleModel
    ds = golden_datasets["synthetic"]
    rng = np.random.default_rng(4)
    n = 60000
    eng = misti_b200.Engine(0)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    m2 = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    m3 = eng.add_model(gid, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    eng.set_data([ds["sfs"]], True)
    p2 = np.zeros((n, 3)); p2[:, 0] = rng.uniform(0, 3, n)
    p3 = np.column_stack([rng.uniform(0, 1.5, n), rng.uniform(0, 1.5, n), rng.uniform(0, 0.3, n)])
    o2 = eng.evaluate(p2, model=m2, flags=15, want=("jafs", "status"))
    o3 = eng.evaluate(p3, model=m3, flags=15, want=("jafs", "status"))
    eng.close()
    checked = 0
    for b in rng.choice(n, 6, replace=False):
        om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], 40, [[2, 5, 12, 0.8, 1]], [], cpfit=True, smooth=True, unfolded=True)
        ref = om.likelihood([float(p2[b, 0])])
        assert o2["status"][b] == 0
        assert relerr(o2["llh"][b, 0], ref) < TOL and relerr(o2["jafs"][b], om.JAFS) < TOL, b
        checked += 1
    for b in np.nonzero(o3["status"] == 0)[0][:6]:
        om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], 40, [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]], [[1, 7, 0.05, 1]], cpfit=True,
                         smooth=True, unfolded=True)
        ref = om.likelihood([float(v) for v in p3[b]])
        assert relerr(o3["llh"][b, 0], ref) < TOL and relerr(o3["jafs"][b], om.JAFS) < TOL, b
        checked += 1
    assert checked == 12
