"""The N > 1 path on CPU: world_size-2 Gloo job exercising the item sharding and the all-gather of the
likelihood rows (misti_b200/parallel.py).  The per-rank evaluator is a stand-in (the CUDA engine needs a
GPU); what is tested is the distributed plumbing: who evaluates what, and that every rank ends up with the
rows in item order."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from misti_b200.parallel import evaluate_sharded, gather_rows, shard_indices, shard_sizes, solve_sharded


def _fake_llh(params, mids):
    # a deterministic function of (params, model id) with R = 3 "data rows"
    base = (params ** 2).sum(axis=1) + (0 if mids is None else 100.0 * np.asarray(mids))
    return np.stack([base, base + 0.5, -base], axis=1)


def _fake_solve(pairs):
    # a deterministic "fit" per (model, row) pair; model 2 has two parameters, the others one (x padded with NaN)
    pairs = np.asarray(pairs)
    n = len(pairs)
    x = np.full((n, 2), np.nan)
    x[:, 0] = pairs[:, 0] + 0.1 * pairs[:, 1]
    x[pairs[:, 0] == 2, 1] = 7.0
    return {"x": x, "llh": -(pairs[:, 0] * 10.0 + pairs[:, 1]), "nfev": 30 + pairs[:, 1], "nit": 15 + pairs[:, 0],
            "success": pairs[:, 1] != 3}


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    params = rng.uniform(0, 3, (B, 2))
    mids = rng.integers(0, 4, B)
    seen = []

    def ev(p, m):
        seen.append(len(p))
        return _fake_llh(p, m)
    full = evaluate_sharded(ev, params, mids)
    ok = np.array_equal(full, _fake_llh(params, mids)) and seen == [shard_sizes(B, world)[rank]]
    # rows of unequal shard sizes (B odd) and a scalar column
    col = gather_rows(np.arange(rank, B, world, dtype=float).reshape(-1, 1), B)
    ok = ok and np.array_equal(col[:, 0], np.arange(B, dtype=float))
    # fits: every rank solves its share of the (model, row) pairs and ends up with all results in pair order
    pairs = [(m, r) for r in range(5) for m in range(3)]
    res = solve_sharded(_fake_solve, pairs)
    ref = _fake_solve(np.array(pairs))
    ok = ok and all(np.array_equal(res[k], ref[k], equal_nan=True) for k in ("x", "llh", "nfev", "nit", "success"))
    ok = ok and np.array_equal(res["model"], np.array(pairs)[:, 0])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharding_covers_every_item_once():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            got = np.concatenate([shard_indices(n, r, world) for r in range(world)]) if n else np.array([])
            assert sorted(got.tolist()) == list(range(n))
            assert shard_sizes(n, world) == [len(shard_indices(n, r, world)) for r in range(world)]


def test_world_size_two_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 37, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_single_process_passthrough():
    params = np.random.default_rng(1).uniform(0, 1, (5, 2))
    assert np.array_equal(evaluate_sharded(_fake_llh, params), _fake_llh(params, None))
    pairs = [(m, r) for r in range(4) for m in range(3)]
    res, ref = solve_sharded(_fake_solve, pairs), _fake_solve(np.array(pairs))
    assert all(np.array_equal(res[k], ref[k], equal_nan=True) for k in ("x", "llh", "nfev", "nit", "success"))
