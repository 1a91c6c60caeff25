"""Sharding of a batch of independent evaluation items over the ranks of a torch.distributed job.

One process per GPU.  Every item (parameter vector x split time x band layout; bootstrap rows ride
along inside an item) is independent, so the only collective of the whole path is the all-gather of
the small likelihood vectors back to every rank.  Items are dealt round-robin
(item i -> rank i mod world) rather than in contiguous blocks because the cost of an item grows with
its split index and, for the optimisers, with its iteration count.
"""
import numpy as np


def shard_indices(n_items, rank, world):
    """Indices of the items rank `rank` evaluates (round-robin)."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def shard_sizes(n_items, world):
    return [(n_items - r + world - 1) // world for r in range(world)]


def gather_rows(local_rows, n_items, group=None, device=None):
    """All-gather per-item rows computed on each rank's shard back into item order on every rank.

    local_rows: array [n_local, K] for the items shard_indices(n_items, rank, world), in that order.
    Returns a numpy array [n_items, K].  Backend-agnostic: NCCL (tensors staged on `device`) or Gloo (CPU).
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local_rows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    local_rows = local_rows.reshape(local_rows.shape[0], -1)
    K = local_rows.shape[1]
    sizes = shard_sizes(n_items, world)
    assert local_rows.shape[0] == sizes[rank], "shard size mismatch"
    pad = max(sizes)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    send = torch.zeros((pad, K), dtype=torch.float64, device=dev)
    send[:sizes[rank]] = torch.from_numpy(local_rows).to(dev)
    recv = torch.empty((world * pad, K), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, pad, K)
    out = np.empty((n_items, K))
    for r in range(world):
        out[r::world] = recv[r, :sizes[r]]
    return out


def evaluate_sharded(evaluate, params, model_ids=None, group=None, device=None):
    """Evaluate a global batch cooperatively.  `evaluate(params_shard, model_ids_shard) -> llh [n_local, R]` runs
    this rank's shard (e.g. a closure over Engine.evaluate); every rank returns the full llh [B, R]."""
    import torch.distributed as dist
    params = np.asarray(params, dtype=np.float64)
    B = params.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    idx = shard_indices(B, rank, world)
    mids = None if model_ids is None else np.asarray(model_ids)[idx]
    local = np.asarray(evaluate(params[idx], mids), dtype=np.float64)
    local = local.reshape(len(idx), -1)
    return gather_rows(local, B, group=group, device=device)


def solve_sharded(solve, pairs, group=None, device=None):
    """Fit a global list of independent (model, data row) pairs cooperatively: rank r fits the pairs
    shard_indices(len(pairs), r, world) with `solve(pairs_shard) -> dict` (e.g. Sweep.solve: arrays x [n, P], llh [n],
    nfev [n], nit [n], success [n] in the order of the shard) and every rank returns the dict for ALL pairs, in pair order.
    The only collective is the all-gather of these few numbers per fit (basin-hopping walkers and bootstrap x split-time
    fits shard the same way)."""
    import torch.distributed as dist
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    K = pairs.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    idx = shard_indices(K, rank, world)
    res = solve(pairs[idx]) if len(idx) else None
    # the width of x is the same on every rank only if every rank knows the widest model: take it from the results
    P_local = 0 if res is None else int(np.asarray(res["x"]).reshape(len(idx), -1).shape[1])
    if world > 1:
        import torch
        t = torch.tensor([P_local], dtype=torch.int64, device=torch.device(device) if device is not None else torch.device("cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        P = int(t.item())
    else:
        P = P_local
    cols = np.full((len(idx), P + 4), np.nan)
    if res is not None:
        cols[:, :P_local] = np.asarray(res["x"], dtype=np.float64).reshape(len(idx), -1)
        for j, k in enumerate(("llh", "nfev", "nit", "success")):
            cols[:, P + j] = np.asarray(res[k], dtype=np.float64)
    full = gather_rows(cols, K, group=group, device=device)
    return {"model": pairs[:, 0], "row": pairs[:, 1], "x": full[:, :P], "llh": full[:, P], "nfev": full[:, P + 1].astype(np.int64),
            "nit": full[:, P + 2].astype(np.int64), "success": full[:, P + 3] != 0}
