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


def gather_rows_device(local, n_items, group=None):
    """The collective of the path, device-resident: `local` is a torch tensor [n_local, K] ON THE DEVICE holding the rows of
    the items shard_indices(n_items, rank, world); returns a device tensor [n_items, K] in item order on every rank (one
    NCCL all-gather over NVLink and one strided device copy; nothing is staged through the host)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_items, world)
    assert local.shape[0] == sizes[rank], "shard size mismatch"
    K = local.shape[1]
    pad = max(sizes)
    if sizes[rank] == pad:
        send = local.contiguous()
    else:
        send = torch.zeros((pad, K), dtype=local.dtype, device=local.device)
        send[:sizes[rank]] = local
    recv = torch.empty((world, pad, K), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(world * pad, K), send, group=group)
    if n_items == world * pad:  # every shard full: item i = recv[i % world, i // world]
        return recv.permute(1, 0, 2).reshape(n_items, K)
    out = torch.empty((n_items, K), dtype=local.dtype, device=local.device)
    for r in range(world):
        out[r::world] = recv[r, :sizes[r]]
    return out


def gather_rows(local_rows, n_items, group=None, device=None):
    """All-gather per-item rows computed on each rank's shard back into item order on every rank.

    local_rows: array [n_local, K] for the items shard_indices(n_items, rank, world), in that order -- a numpy array
    (returned as numpy [n_items, K]; with NCCL it crosses to `device` for the collective) or a torch tensor on the device
    (returned as a device tensor, see gather_rows_device).  Backend-agnostic: NCCL or Gloo (CPU)."""
    import torch
    import torch.distributed as dist
    if isinstance(local_rows, torch.Tensor):
        return gather_rows_device(local_rows.reshape(local_rows.shape[0], -1), n_items, group=group)
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local_rows)
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    local_rows = local_rows.reshape(local_rows.shape[0], -1)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    return gather_rows_device(torch.from_numpy(local_rows).to(dev), n_items, group=group).cpu().numpy()


class ShardedEvaluator:
    """The batched objective over the ranks of a job, device-resident: every rank evaluates its shard of a global batch on
    its own GPU (Engine.evaluate_device on its torch stream) and ONE all-gather brings the likelihood rows to every rank
     -- the call an optimiser that lives on every rank makes per step.  Buffers are allocated once."""

    def __init__(self, engine, device, n_local, P, group=None, want_jafs=False, zero_copy=True):
        import torch
        import torch.distributed as dist
        self.engine, self.group, self.device, self.zero_copy = engine, group, torch.device(device), zero_copy
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.n_local, self.P, self.R = int(n_local), int(P), engine.R
        self.params_d = torch.empty((n_local, max(P, 1)), dtype=torch.float64, device=self.device)
        self.llh_d = torch.empty((n_local, self.R), dtype=torch.float64, device=self.device)
        self.status_d = torch.empty((n_local,), dtype=torch.int32, device=self.device)
        self.jafs_d = torch.empty((n_local, 7), dtype=torch.float64, device=self.device) if want_jafs else None
        self.llh_all_h = torch.empty((self.world * n_local, self.R), dtype=torch.float64).pin_memory()
        self.status_h = torch.empty((n_local,), dtype=torch.int32).pin_memory()
        self.jafs_h = torch.empty((n_local, 7), dtype=torch.float64).pin_memory() if want_jafs else None

    def evaluate_device(self, params_d, model, flags):
        """shard on the device -> global likelihood rows on the device [world * n_local, R], item order"""
        self.engine.evaluate_device(self.n_local, self.P, params_d.data_ptr() if self.P else None, self.llh_d.data_ptr(), model=model,
                                    flags=flags, status_ptr=self.status_d.data_ptr(),
                                    jafs_ptr=self.jafs_d.data_ptr() if self.jafs_d is not None else None)
        return gather_rows_device(self.llh_d, self.world * self.n_local, group=self.group)

    def evaluate(self, params_h, model, flags):
        """host shard (pinned tensor [n_local, P]) -> global likelihood rows on the host (pinned), + this shard's status and
        spectra.  H2D of the parameters, the two kernels, the all-gather and the D2H of the results are queued on the
        current torch stream; returns after they have completed.
        One rank (no collective): the kernels work on the pinned host buffers DIRECTLY -- pinned memory is device-accessible
        (unified addressing), so the correction kernel reads its parameters and the JSFS kernel writes likelihoods, spectra
        and status across PCIe while it computes, and no separate copy trails the step."""
        import torch
        if self.world == 1 and self.zero_copy and params_h.is_pinned():
            self.engine.evaluate_device(self.n_local, self.P, params_h.data_ptr() if self.P else None, self.llh_all_h.data_ptr(),
                                        model=model, flags=flags, status_ptr=self.status_h.data_ptr(),
                                        jafs_ptr=self.jafs_h.data_ptr() if self.jafs_h is not None else None)
            torch.cuda.current_stream(self.device).synchronize()
            return self.llh_all_h, self.status_h, self.jafs_h
        if self.zero_copy and params_h.is_pinned():
            # several ranks: parameters, spectra and status cross PCIe from / to the pinned buffers inside the kernels as
            # above; the likelihood rows go through the all-gather on the device and are copied out
            self.engine.evaluate_device(self.n_local, self.P, params_h.data_ptr() if self.P else None, self.llh_d.data_ptr(), model=model,
                                        flags=flags, status_ptr=self.status_h.data_ptr(),
                                        jafs_ptr=self.jafs_h.data_ptr() if self.jafs_h is not None else None)
            llh_all = gather_rows_device(self.llh_d, self.world * self.n_local, group=self.group)
            self.llh_all_h.copy_(llh_all, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return self.llh_all_h, self.status_h, self.jafs_h
        self.params_d.copy_(params_h, non_blocking=True)
        llh_all = self.evaluate_device(self.params_d, model, flags)
        self.llh_all_h.copy_(llh_all, non_blocking=True)
        self.status_h.copy_(self.status_d, non_blocking=True)
        if self.jafs_h is not None:
            self.jafs_h.copy_(self.jafs_d, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.llh_all_h, self.status_h, self.jafs_h


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
