#!/usr/bin/env python3
"""The fused likelihood stage with bootstrap rows: B items scored against R data rows in the JSFS kernel
(llh[b, r] = const_r + sum_i d_ri log p_bi).  Device-resident buffers; reports the kernel time with R = 1 and R = 1001
and the HBM rate of the difference (8 bytes written per (item, row) pair; the 64-byte data rows are read from L2).
Run under gpurun; prints JSON."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import misti_b200  # noqa: E402
from misti_b200 import io as mio  # noqa: E402


def main():
    import torch
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = torch.device("cuda", 0)
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]["synthetic"]
    bs = mio.read_jafs(os.path.join(ROOT, "data", "synthetic", "bs.sfs")).jafs
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng = misti_b200.Engine(0, stream=stream.cuda_stream)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, 40, 0, bands=[(1, 5, 12, 0.8, 0)])
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED
    p = torch.from_numpy(np.random.default_rng(1234).uniform(0, 5, (B, 1))).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = {"B": B}
    k2 = {}
    for R in (1, len(bs)):
        eng.set_data(bs[:R], True)
        llh = torch.empty((B, R), dtype=torch.float64, device=dev)
        ms = []
        for it in range(8):
            flush.zero_()
            eng.evaluate_device(B, 1, p.data_ptr(), llh.data_ptr(), model=mid, flags=flags)
            a, b = eng.last_kernel_ms()
            if it >= 3:
                ms.append(b)
        k2[R] = float(np.median(ms))
        out["R=%d" % R] = {"jsfs_kernel_ms": k2[R], "llh_bytes_written": B * R * 8, "finite": bool(torch.isfinite(llh).all().item())}
    R = len(bs)
    extra_ms = k2[R] - k2[1]
    out["likelihood_stage"] = {"pairs": B * R, "extra_ms": extra_ms, "hbm_write_GBps": B * (R - 1) * 8 / (extra_ms * 1e-3) / 1e9,
                               "pairs_per_s": B * (R - 1) / (extra_ms * 1e-3),
                               "algorithmic_bytes_per_pair": 72, "note": "SURVEY 8d: 64 B data row (L2-resident, 64 KB in all) + 8 B written per pair"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
