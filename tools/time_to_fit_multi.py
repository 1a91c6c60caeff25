#!/usr/bin/env python3
"""Time-to-fit of BASELINE configs 3 and 5b on N GPUs of one box (one process per GPU under torchrun; also runs as a
single process).  The fits are independent: the (bootstrap row x split time) pairs of config 5b and the basin-hopping
walkers of config 3 are dealt over the ranks, and one all-gather (NCCL) brings the few numbers per fit back
(misti_b200.parallel.solve_sharded / gather_rows).  Rank 0 prints one JSON object.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/time_to_fit_multi.py [walkers] [hops]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402
from misti_b200 import io as mio  # noqa: E402
from misti_b200.optim import basinhopping_batch  # noqa: E402
from misti_b200.parallel import gather_rows, shard_indices  # noqa: E402
from misti_b200.sweep import Sweep  # noqa: E402

DATA = os.path.join(ROOT, "data", "synthetic")


def main():
    walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    niter = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    units = mio.Units.from_file(os.path.join(DATA, "setunits.txt"))
    inp = mio.read_psmc(os.path.join(DATA, "m1.psmc"), os.path.join(DATA, "m2.psmc"), 0, -1, units)
    data = mio.column_sums(mio.read_jafs(os.path.join(DATA, "m.sfs")).jafs)
    bs = mio.read_jafs(os.path.join(DATA, "bs.sfs")).jafs
    eng = misti_b200.Engine(local)
    out = {"n_gpus": world}

    # warm-up: context, kernel variants, NCCL
    sw = Sweep(inp.times, inp.lambdas, bs[:8], unfolded=True, cpfit=True, smooth=True, engine=eng)
    sw.add_model(40, [[1, 4, 40, 3, 1]])
    sw.solve_distributed(device=dev)

    # config 5b: 1001 rows x split times 36..44, one Nelder-Mead fit each
    barrier()
    t = time.perf_counter()
    sw = Sweep(inp.times, inp.lambdas, bs, unfolded=True, cpfit=True, smooth=True, engine=eng)
    for st in range(36, 45):
        sw.add_model(st, [[1, 4, st, 3, 1]])
    r = sw.solve_distributed(tol=1e-4, device=dev)
    barrier()
    dt = time.perf_counter() - t
    out["config5b_band_to_split_fits"] = {"s": dt, "fits": len(r["llh"]), "converged": int(r["success"].sum()),
                                          "scipy_nfev_total": int(r["nfev"].sum()), "checksum_llh": float(np.sum(r["llh"]))}

    # config 3: basin-hopping walkers (two bands + pulse), local searches on the device, walkers dealt over the ranks
    barrier()
    t = time.perf_counter()
    sw = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
    m = sw.add_model(40, [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]], [[1, 7, 0.05, 1]])
    rng = np.random.default_rng(2024)
    x0 = np.column_stack([rng.uniform(0, 5, walkers), rng.uniform(0, 5, walkers), rng.uniform(0, 0.5, walkers)])
    mine = shard_indices(walkers, rank, world)
    mids = np.full(len(mine), sw.models[m]["id"], dtype=np.int32)
    res = basinhopping_batch(None, x0[mine], niter=niter, T=0.5, seeds=[2024 + int(w) for w in mine],
                             local_solver=lambda xs: eng.nelder_mead(xs, mids, np.zeros(len(mine), dtype=np.int32), flags=sw.flags))
    rows = gather_rows(np.column_stack([res["x"], res["fun"], res["nfev"]]), walkers, device=dev)
    barrier()
    dt = time.perf_counter() - t
    best = int(np.argmin(rows[:, 3]))
    out["config3_basinhopping"] = {"s": dt, "walkers": walkers, "niter": niter, "best_x": rows[best, :3].tolist(),
                                   "best_llh": float(-rows[best, 3]), "scipy_nfev_total": int(rows[:, 4].sum())}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
