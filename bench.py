#!/usr/bin/env python3
"""bench.py -- batched expected-JSFS + composite-logL evaluations per second, and time to fit (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of B synthetic parameter vectors per GPU (BASELINE config 2: split
index 40, unfolded SFS, one optimised migration band `-mi 2 5 12 0.8 1`, `--cpfit`, m ~ U(0,5), numpy
default_rng(1234 + rank)).  Prints ONE JSON line.

  value        evaluations/s, whole job, inputs resident in HBM, CUDA-event timed per step.  At N > 1 every timed step
               ends with the path's one collective, the all-gather of the likelihood vectors to every rank
               (misti_b200.parallel.gather_rows_device, device-resident) -- `value_no_collective` is the same without it
  e2e          the same metric through the public host-buffer API (misti_b200.parallel.ShardedEvaluator): pinned host buffers,
               parameters host -> device, the kernels, (N > 1: the all-gather,) llh + expected JSFS + status device -> host
               inside the timed region.  The pinned buffers are device-accessible, so the kernels read / write them across
               PCIe themselves while they compute (--no-zero-copy: staged cudaMemcpyAsync instead, 8 % slower)
  roofline     the dominant kernel (misti_correct_kernel) against the measured FP64 peak of this pool's B200: FLOPs it
               executed (SASS counts of the committed ncu capture) / its CUDA-event time; the second kernel and the
               dense-equivalent figure of SURVEY.md 8d are listed beside it under their own keys
  batch_sweep  evaluations/s at B = 2, 64, 4096, 65536
  config_sweep kernel time of one batch for the layouts of BASELINE configs 1, 3, 4 and config 2 in the reference's default mode
  likelihood_stage  config 5's bootstrap stage: 65 536 items x 1 001 data rows, HBM write rate against the measured copy rate
  time_to_fit  the second half of the metric, device-timed, fits sharded over the N ranks (strong scaling):
               config 2 (one Nelder-Mead fit), config 5b (9 009 fits = 1 001 bootstrap rows x 9 split times), config 3
               (basin-hopping walkers, 1 024 per GPU)
  cpu_baseline the reference's own CPU implementation (oracle/_ref, the unmodified Genomics-HSE/MiSTI staged by
               oracle/make_ref.py; kind "reference") -- or, where that is absent, the oracle port (kind "port") -- on
               the host cores, bounded sample of the same parameter vectors (rank 0, N = 1 only)

--impl reference times that CPU implementation alone, one worker process per host core (BLAS threads 1 as MiSTI.py:23-25).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_k, "1")  # as MiSTI.py:23-25

METRIC = "expected-JSFS+logL evals/sec"
_emit = print  # replaced in main() by a writer to the process's original stdout
SPLIT_T, BAND = 40, [2, 5, 12, 0.8, 1]
NUM_T = 127
# dense formulation (Pade-13 + Van Loan, zero squarings) per evaluation
F_DENSE = SPLIT_T * (12 + 8.0 / 3) * 45 ** 3 + (NUM_T - SPLIT_T) * (12 + 8.0 / 3) * 9 ** 3
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")


def _first_existing(*names):
    for n in names:
        p = os.path.join(ROOT, "profiles", n)
        if os.path.exists(p):
            return p
    return None


EXEC_FLOPS_FILE = _first_existing("r03_executed_flops.json", "r02_executed_flops.json", "r01_executed_flops.json")


def load_dataset():
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        return json.load(f)["datasets"]["synthetic"]


def make_params(B, rank):
    import numpy as np
    return np.random.default_rng(1234 + rank).uniform(0.0, 5.0, (B, 1))


def workload_name(B):
    return ("config2: m1.psmc m2.psmc m.sfs st=40 -uf -mi 2 5 12 0.8 1 --cpfit (numT=127), batched objective, "
            "B=%d vectors/GPU, m~U(0,5)" % B)


def bench_config(B):
    """the same dict in both arms (the driver compares them)"""
    return {"workload": workload_name(B), "l2": "256 MiB buffer rewritten between timed iterations",
            "parallelism": "independent items sharded across ranks; all_gather of llh inside every timed step at N > 1"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) when staged, else the oracle port; one worker process per core
# ------------------------------------------------------------------------------------------------
_worker_model = None


def cpu_kind():
    from oracle import ref_loader
    return "reference" if ref_loader.available() else "port"


def _cpu_model(mi=(BAND,), pu=(), st=SPLIT_T):
    from oracle import ref_loader
    ds = load_dataset()
    if ref_loader.available():
        M = ref_loader.make_model(ds["times"], ds["lambdas"], ds["sfs"], st, mi, pu, cpfit=True, smooth=True, unfolded=True)
        return lambda x: float(ref_loader.quiet(M.JAFSLikelihood, list(x))), M
    from oracle.misti_oracle import OracleModel
    M = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], st, [list(m) for m in mi], [list(p) for p in pu], cpfit=True, smooth=True,
                    unfolded=True)
    return lambda x: float(M.likelihood(list(x))), M


def _worker_eval(m):
    global _worker_model
    if _worker_model is None:
        _worker_model = _cpu_model()[0]
    return _worker_model([m])


def cpu_path_text():
    import numpy
    import scipy
    if cpu_kind() == "reference":
        return ("oracle/_ref/misti_reference.zip: Genomics-HSE/MiSTI unmodified (MigrationInference.JAFSLikelihood), numpy %s / "
                "scipy %s" % (numpy.__version__, scipy.__version__))
    return "oracle/misti_oracle.py (numpy %s / scipy %s port of the reference path; oracle/_ref is not staged)" % (
        numpy.__version__, scipy.__version__)


def cpu_rate(n_evals, cores, rank=0):
    """evals/s of the CPU implementation on `cores` worker processes over n_evals parameter vectors."""
    import multiprocessing as mp
    ms = [float(v) for v in make_params(n_evals, rank)[:, 0]]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_worker_eval, ms[:cores])  # start-up + imports outside the timed region
        t0 = time.perf_counter()
        pool.map(_worker_eval, ms, chunksize=1)
        dt = time.perf_counter() - t0
    return n_evals / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    per_step = (2 if kind == "reference" else 16) * cores  # a step stays well under a second per core
    import multiprocessing as mp
    ms_all = [float(v) for v in make_params(per_step * (args.steps + args.warmup), 0)[:, 0]]
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(cores) as pool:
        pool.map(_worker_eval, ms_all[:cores])
        for s in range(args.warmup + args.steps):
            chunk = ms_all[s * per_step:(s + 1) * per_step]
            t0 = time.perf_counter()
            pool.map(_worker_eval, chunk, chunksize=1)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    value = per_step * args.steps / total
    sample = "%d evaluations per step (same parameter distribution), %d worker processes, BLAS threads 1" % (per_step, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(args.batch), "cpu_path": cpu_path_text(),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.monotonic()] + [c.strip() for c in line.split(",")])

    def stop(self, windows):
        """median SM clock over the samples taken inside the timed windows [(t0, t1), ...]"""
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        rows = [r[1:] for r in self.rows if any(t0 <= r[0] <= t1 + 0.02 for t0, t1 in windows)]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_to_fit(args, eng, dev, world, rank, barrier, stream):
    """Device-timed fits, sharded over the ranks (strong scaling for config 5b, 1 024 walkers per GPU for config 3)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import misti_b200
    from misti_b200 import io as mio
    from misti_b200.parallel import gather_rows, shard_indices
    from misti_b200.sweep import Sweep
    data_dir = os.path.join(ROOT, "data", "synthetic")
    units = mio.Units.from_file(os.path.join(data_dir, "setunits.txt"))
    inp = mio.read_psmc(os.path.join(data_dir, "m1.psmc"), os.path.join(data_dir, "m2.psmc"), 0, -1, units)
    data = mio.column_sums(mio.read_jafs(os.path.join(data_dir, "m.sfs")).jafs)
    bs = mio.read_jafs(os.path.join(data_dir, "bs.sfs")).jafs
    out = {}

    def timed(fn, reps=2):
        """fn() run `reps` times (the first pays graph capture and buffer growth); device time of the last, max over ranks"""
        res, ms = None, 0.0
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            res = fn()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            wall = time.perf_counter() - t0
        t = torch.tensor([ms, 1e3 * wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t[0].item()) * 1e-3, float(t[1].item()) * 1e-3

    # ---- config 5b: 1 001 rows x split times 36..44, `-uf -mi 1 4 st 3 1 --cpfit`: 9 009 Nelder-Mead fits, dealt over the ranks
    sts = list(range(36, 45))
    sw = Sweep(inp.times, inp.lambdas, bs, unfolded=True, cpfit=True, smooth=True, engine=eng)
    for st in sts:
        sw.add_model(st, [[1, 4, st, 3, 1]])
    pairs = np.array([(m, r) for r in range(len(bs)) for m in range(len(sts))], dtype=np.int64)
    mine = pairs[shard_indices(len(pairs), rank, world)]

    def fits_5b():
        r = sw.solve(pairs=mine, tol=1e-4)
        cols = torch.from_numpy(np.column_stack([r["x"][:, 0], r["llh"], r["nfev"].astype(np.float64)])).to(dev)
        full = gather_rows(cols, len(pairs))  # x, llh, nfev of every fit on every rank: the collective of the fit path
        return r, full
    (r5, full5), s5, w5 = timed(fits_5b)
    full5 = full5.cpu().numpy()
    out["config5b"] = {"what": "9 009 Nelder-Mead fits (1 001 bootstrap rows x split times 36..44, -uf -mi 1 4 st 3 1 --cpfit, tol 1e-4), "
                               "fits dealt over the ranks, x / llh / nfev all-gathered",
                       "scaling": "strong", "fits": int(len(pairs)), "device_s": s5, "wall_s": w5, "fits_per_s": len(pairs) / s5,
                       "rounds_of_launches_rank0": int(r5["launches"]), "device_evaluations_rank0": int(r5["evaluations"]),
                       "scipy_nfev_total": int(full5[:, 2].sum()), "checksum_llh": float(full5[:, 1].sum()),
                       "converged_rank0": int(r5["success"].sum()), "limiter": "rounds x latency of one item's serial correction chain "
                       "(36 trust-region intervals per item; the first rounds start in the run-away regime m = 3)"}

    # ---- config 3: two bands + pulse, basin-hopping walkers (1 024 per GPU), niter hops, T = 0.5, stepsize = 0.5, seed 2024
    sw3 = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
    m3 = sw3.add_model(40, [[1, 2, 10, 0.3, 1], [2, 5, 12, 0.8, 1]], [[1, 7, 0.05, 1]])
    W = args.walkers
    rng = np.random.default_rng(2024 + rank)
    x0 = np.column_stack([rng.uniform(0, 5, W), rng.uniform(0, 5, W), rng.uniform(0, 0.5, W)])
    mids = np.full(W, sw3.models[m3]["id"], dtype=np.int32)
    seeds = [2024 + rank * W + w for w in range(W)]

    def walkers_3():
        r = eng.basinhopping(x0, mids, np.zeros(W, dtype=np.int32), seeds=seeds, flags=sw3.flags, niter=args.fit_niter, T=0.5, stepsize=0.5)
        best = int(np.argmin(r["fun"]))
        row = torch.tensor([[r["fun"][best]] + r["x"][best].tolist() + [float(r["nfev"].sum()), float(r["evaluations"])]],
                           dtype=torch.float64, device=dev)
        return r, gather_rows(row, world)  # every rank learns every rank's best walker
    (r3, best3), s3, w3 = timed(walkers_3, reps=1 if args.fit_niter >= 50 else 2)
    best3 = best3.cpu().numpy()
    k = int(np.argmin(best3[:, 0]))
    out["config3"] = {"what": "basin-hopping (-mi 1 2 10 0.3 1 -mi 2 5 12 0.8 1 -pu 1 7 0.05 1 --cpfit), T = 0.5, stepsize = 0.5, seeds 2024.., "
                              "%d walkers per GPU x %d hops, walkers advance independently on the device" % (W, args.fit_niter),
                      "scaling": "weak", "walkers_total": W * world, "niter": args.fit_niter, "device_s": s3, "wall_s": w3,
                      "scipy_nfev_total": int(best3[:, 4].sum()), "device_evaluations_total": int(best3[:, 5].sum()),
                      "scipy_evals_per_s": float(best3[:, 4].sum()) / s3, "rounds_of_launches_rank0": int(r3["launches"]),
                      "best_llh": float(-best3[k, 0]), "best_x": [float(v) for v in best3[k, 1:4]],
                      "limiter": "rounds x latency of one evaluation (a walker's hops and iterations are sequential)"}
    if rank != 0:
        return None

    # ---- config 2: ONE Nelder-Mead fit (latency of a serial fit), next to the CPU implementation on one core
    sw2 = Sweep(inp.times, inp.lambdas, [data], unfolded=True, cpfit=True, engine=eng)
    sw2.add_model(SPLIT_T, [BAND])
    r2, s2, w2 = None, None, None
    for _ in range(3):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        r2 = sw2.solve(tol=1e-4)
        w2 = time.perf_counter() - t0
    out["config2"] = {"what": "one Nelder-Mead fit of the optimised band (tol 1e-4, start 0.8)", "wall_s": w2, "x": r2["x"][0][:1].tolist(),
                      "llh": float(r2["llh"][0]), "nfev": int(r2["nfev"][0]), "rounds_of_launches": int(r2["launches"])}
    if world == 1 and not args.skip_cpu:
        from scipy import optimize
        f, _ = _cpu_model()
        t0 = time.perf_counter()
        ref = optimize.minimize(lambda x: -f(x), [BAND[3]], method="Nelder-Mead", options={"xatol": 1e-4, "fatol": 1e-4, "maxiter": 1000})
        cpu_s = time.perf_counter() - t0
        out["config2"].update({"cpu_s": cpu_s, "cpu_kind": cpu_kind(), "cpu_cores": 1, "cpu_x": [float(v) for v in ref.x],
                               "cpu_llh": float(-ref.fun), "cpu_nfev": int(ref.nfev)})
        per_eval = cpu_s / max(1, ref.nfev)
        out["config5b"]["cpu_extrapolated_s_1core"] = per_eval * out["config5b"]["scipy_nfev_total"]
        out["config3"]["cpu_extrapolated_s_1core"] = per_eval * out["config3"]["scipy_nfev_total"]
        out["cpu_note"] = ("cpu_extrapolated_s_1core = the CPU implementation's measured seconds per evaluation in the config-2 fit "
                           "x scipy's evaluation count of the fits (its models cost the CPU at least as much per evaluation)")
    return out


def config_sweep(local, stream, B, rank):
    """Kernel time of one batch of B evaluations for the layouts of the other BASELINE configurations (an engine of its own, so
    that the bench's model stays registered): config 1 (no migration, reference's default mode, the split times 30..60 interleaved),
    config 3 (two bands + pulse, --cpfit), config 4 (ancient second genome: sampling date 12, --hetloss rates, unfolded, a band from
    the sampling date, every split time 30..60 interleaved) and config 2 in the reference's default mode."""
    import json
    import numpy as np
    import misti_b200
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        dss = json.load(f)["datasets"]
    ds, da = dss["synthetic"], dss["synthetic_ancient"]
    eng = misti_b200.Engine(local, stream=stream.cuda_stream)
    g = eng.add_grid(ds["times"], ds["lambdas"])
    ga = eng.add_grid(da["times"], da["lambdas"])
    sd = int(da["sampleDate"])
    rng = np.random.default_rng(77 + rank)
    grid1 = np.array([eng.add_model(g, st, 0) for st in range(30, 61)], dtype=np.int32)
    m2 = eng.add_model(g, SPLIT_T, 0, bands=[(BAND[0] - 1, BAND[1], BAND[2], BAND[3], 0)])
    m3 = eng.add_model(g, 40, 0, bands=[(0, 2, 10, 0.3, 0), (1, 5, 12, 0.8, 1)], pulses=[(0, 7, 0.05, 2)])
    grid4 = np.array([eng.add_model(ga, st, sd, bands=[(1, sd, sd + 8, 0.5, 0)]) for st in range(30, 61)], dtype=np.int32)
    F = misti_b200
    cases = [("config1_no_migration_default_mode_split_grid", ds, dict(model_ids=grid1[np.arange(B) % 31]), np.zeros((B, 1)), F.FLAG_CORRECT | F.FLAG_SMOOTH | F.FLAG_UNFOLDED),
             ("config2_default_mode", ds, dict(model=m2), rng.uniform(0, 5, (B, 1)), F.FLAG_CORRECT | F.FLAG_SMOOTH | F.FLAG_UNFOLDED),
             ("config3_two_bands_pulse_cpfit", ds, dict(model=m3), np.column_stack([rng.uniform(0, 5, B), rng.uniform(0, 5, B), rng.uniform(0, 0.5, B)]),
              F.FLAG_CORRECT | F.FLAG_CPFIT | F.FLAG_SMOOTH | F.FLAG_UNFOLDED),
             ("config4_ancient_sample_split_grid_cpfit", da, dict(model_ids=grid4[np.arange(B) % 31]), rng.uniform(0, 3, (B, 1)),
              F.FLAG_CORRECT | F.FLAG_CPFIT | F.FLAG_SMOOTH | F.FLAG_UNFOLDED)]
    out = {}
    for name, d, kw, par, flags in cases:
        par = np.ascontiguousarray(np.pad(par, ((0, 0), (0, 3 - par.shape[1]))))  # P = the largest model's parameter count
        eng.set_data([d["sfs"]], True)
        ts = []
        for _ in range(5):
            o = eng.evaluate(par, flags=flags, want=("status",), **kw)
            ts.append(eng.last_kernel_ms())
        k1, k2 = float(np.median([a for a, _ in ts])), float(np.median([b for _, b in ts]))
        out[name] = {"B": B, "correction_kernels_ms": k1, "jsfs_likelihood_kernels_ms": k2, "evals_per_s_device": B / ((k1 + k2) * 1e-3),
                     "ok_fraction": float((o["status"] == 0).mean())}
    eng.close()
    return out


def likelihood_stage(local, stream, B):
    """BASELINE config 5's bootstrap stage: B items scored against the 1 001 rows of data/synthetic/bs.sfs in one call
    (llh[b, r] = const_r + sum_i d_ri log p_bi: 8 bytes written per (item, row) pair) -- the HBM-bound part of the path.
    Device-resident buffers; the stage's time = the JSFS + likelihood kernels with 1 001 rows minus the same launch with one row."""
    import json
    import numpy as np
    import torch
    import misti_b200
    from misti_b200 import io as mio
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        ds = json.load(f)["datasets"]["synthetic"]
    bs = mio.read_jafs(os.path.join(ROOT, "data", "synthetic", "bs.sfs")).jafs
    dev = torch.device("cuda", local)
    eng = misti_b200.Engine(local, stream=stream.cuda_stream)
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, SPLIT_T, 0, bands=[(BAND[0] - 1, BAND[1], BAND[2], BAND[3], 0)])
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED
    p = torch.from_numpy(np.random.default_rng(99).uniform(0, 5, (B, 1))).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    k2 = {}
    for R in (1, len(bs)):
        eng.set_data(bs[:R], True)
        llh = torch.empty((B, R), dtype=torch.float64, device=dev)
        ms = []
        for it in range(7):
            flush.zero_()
            eng.evaluate_device(B, 1, p.data_ptr(), llh.data_ptr(), model=mid, flags=flags)
            if it >= 2:
                ms.append(eng.last_kernel_ms()[1])
        k2[R] = float(np.median(ms))
        finite = bool(torch.isfinite(llh).all().item())
        del llh
    eng.close()
    R = len(bs)
    extra = k2[R] - k2[1]
    peak = _measured_hbm()
    gbs = B * (R - 1) * 8 / (extra * 1e-3) / 1e9
    return {"items": B, "rows": R, "pairs": B * R, "jsfs_likelihood_kernels_ms": {"1_row": k2[1], "%d_rows" % R: k2[R]}, "stage_ms": extra,
            "pairs_per_s": B * (R - 1) / (extra * 1e-3), "hbm_write_gbs": gbs, "hbm_peak_gbs": peak, "frac_of_hbm_peak": gbs / peak if peak else None,
            "finite": finite, "bound": "hbm (8 B written per pair; the 64-byte data rows stay in L2)"}


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import misti_b200
    from misti_b200.parallel import ShardedEvaluator, gather_rows_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not land on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(dev)  # a real (non-NULL) stream shared by torch events, NCCL and the engine's launches
    torch.cuda.set_stream(stream)
    eng = misti_b200.Engine(local, stream=stream.cuda_stream)

    ds = load_dataset()
    gid = eng.add_grid(ds["times"], ds["lambdas"])
    mid = eng.add_model(gid, SPLIT_T, 0, bands=[(BAND[0] - 1, BAND[1], BAND[2], BAND[3], 0)])
    eng.set_data([ds["sfs"]], True)
    flags = misti_b200.FLAG_CORRECT | misti_b200.FLAG_CPFIT | misti_b200.FLAG_SMOOTH | misti_b200.FLAG_UNFOLDED

    B = args.batch
    params_h = torch.from_numpy(make_params(B, rank)).pin_memory()
    params_d = params_h.to(dev)
    llh_d = torch.empty((B, 1), dtype=torch.float64, device=dev)
    jafs_d = torch.empty((B, 7), dtype=torch.float64, device=dev)
    status_d = torch.empty((B,), dtype=torch.int32, device=dev)
    terms_d = torch.empty((B,), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(n=B, collective=True):
        eng.evaluate_device(n, 1, params_d.data_ptr(), llh_d.data_ptr(), model=mid, flags=flags, jafs_ptr=jafs_d.data_ptr(),
                            status_ptr=status_d.data_ptr(), terms_ptr=terms_d.data_ptr())
        if world > 1 and collective:
            return gather_rows_device(llh_d[:n], world * n)  # the path's one collective: llh of every item on every rank
        return llh_d

    def timed_steps(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        k1_ms, k2_ms = [], []
        barrier()
        t0 = time.monotonic()
        for s in range(steps):
            flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
            ev[s][0].record(stream)
            fn()
            ev[s][1].record(stream)
            a, b = eng.last_kernel_ms()  # synchronises with this step
            k1_ms.append(a)
            k2_ms.append(b)
        barrier()
        return sum(e0.elapsed_time(e1) for e0, e1 in ev), k1_ms, k2_ms, (t0, time.monotonic())

    # ---- device-resident throughput (with the collective at N > 1) -----------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    windows = []
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    launches0 = eng.launch_count()
    total_ms, k1_ms, k2_ms, win = timed_steps(step_device, args.steps)
    windows.append(win)
    launches = eng.launch_count() - launches0
    nocoll_ms = total_ms
    if world > 1:
        nocoll_ms, _, _, win = timed_steps(lambda: step_device(collective=False), args.steps)
        windows.append(win)

    # ---- end to end through the public host-buffer API -----------------------------------------
    shard = ShardedEvaluator(eng, dev, B, 1, want_jafs=True, zero_copy=not args.no_zero_copy)
    for _ in range(max(1, args.warmup)):
        shard.evaluate(params_h, mid, flags)
    e2e_ms, _, _, win = timed_steps(lambda: shard.evaluate(params_h, mid, flags), args.steps)
    windows.append(win)
    assert torch.equal(shard.llh_all_h[rank::world] if world > 1 else shard.llh_all_h, llh_d.cpu())  # same numbers both ways

    ok_frac = float((status_d == 0).float().mean().item())
    terms_mean = float(terms_d.double().mean().item())

    # ---- throughput vs batch size (B = 2, 64, 4096, 65536) -----------------
    sweep = []
    for n in (2, 64, 4096, 65536):
        if n > B:
            continue
        for _ in range(3):
            step_device(n)
        ms, a, b, win = timed_steps(lambda n=n: step_device(n), 10)
        windows.append(win)
        sweep.append({"B_per_gpu": n, "ms_per_step": ms / 10, "evals_per_s": world * n * 10 / (ms * 1e-3),
                      "kernel_ms": {"misti_correct_kernel": sum(a) / len(a), "misti_jsfs_kernel": sum(b) / len(b)}})

    # ---- the other BASELINE configurations at the same batch size (device time of the kernels) ----
    cfgs = None if args.skip_fits else config_sweep(local, stream, min(B, 65536), rank)
    lstage = None if args.skip_fits else likelihood_stage(local, stream, min(B, 65536))

    # ---- time to fit ---------------------------------------------------------------------------
    t_fit0 = time.monotonic()
    ttf = None if args.skip_fits else time_to_fit(args, eng, dev, world, rank, barrier, stream)
    windows.append((t_fit0, time.monotonic()))
    clocks = sampler.stop(windows) if rank == 0 else None

    t = torch.tensor([total_ms, e2e_ms, nocoll_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, nocoll_ms = (float(v) for v in t.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    k1, k2 = sum(k1_ms) / len(k1_ms), sum(k2_ms) / len(k2_ms)
    with open(FP64_PEAK_FILE) as f:
        peaks = json.load(f)
    peak = peaks["dfma_tflops"]  # the executed kernels are DFMA code
    ex = {}
    if EXEC_FLOPS_FILE:
        with open(EXEC_FLOPS_FILE) as f:
            ex = json.load(f)
    k1_flops = ex.get("misti_correct_kernel", {}).get("flops_per_item")
    if k1_flops is not None:  # the cpfit post-split pass of a large batch runs as a kernel of its own, inside the same pair of events
        post = "misti_post_split_quad_kernel" if (B > 6144 and "misti_post_split_quad_kernel" in ex and os.environ.get("MISTI_POST_QUAD", "1") != "0") \
            else "misti_post_split_kernel"  # plain batches above 6 144 items: four lanes per item
        k1_flops += ex.get(post, {}).get("flops_per_item", 0.0)
    pair_kernel = B > 6144 and os.environ.get("MISTI_JSFS_PAIR", "1") != "0"  # large batches: two lanes per item (csrc/misti_pair.cuh)
    k2_name = "misti_jsfs_pair_kernel" if pair_kernel and "misti_jsfs_pair_kernel" in ex else "misti_jsfs_kernel"
    k2_flops = ex.get(k2_name, {}).get("flops_per_item")
    src = os.path.relpath(EXEC_FLOPS_FILE, ROOT) if EXEC_FLOPS_FILE else None

    def kernel_entry(name, ms, flops, bound):
        tf = None if flops is None else flops * B / (ms * 1e-3) / 1e12
        return {"kernel": name, "ms": ms, "flops_per_eval_executed": flops, "tflops": tf, "frac_of_fp64_peak": None if tf is None else tf / peak,
                "bound": bound}
    dom_is_k1 = k1 >= k2
    kernels = {"misti_correct_kernel": kernel_entry("misti_correct_kernel (+ misti_post_split_quad_kernel)", k1, k1_flops,
                                                    "latency of one thread's serial FP64 chain (one block of 14 warps per SM, 3.5 warps per scheduler; ncu r03: issue slots 36 %, "
                                                    "FP64 pipe 28 %, stalls: long scoreboard 29 % (thread-local stack), fixed latency 26 %, interval barrier 10 %, no "
                                                    "instruction 7 % (20 % before the SM's warps were launched as one block and kept together: 78 KB of hot code)); the "
                                                    "post-split kernel is FP64-bound (issue slots 65 %, FP64 pipe 56 %)"),
               "misti_jsfs_kernel": kernel_entry(
                   k2_name + (" (+ misti_jsfs_kernel over an empty redo list, misti_stiff_kernel with nothing parked)" if k2_name != "misti_jsfs_kernel"
                              else " (+ misti_stiff_kernel, nothing parked)"), k2, k2_flops,
                   "FP64 issue with two warps per scheduler (ncu r03: FP64 pipe 48 %, issue slots 44 %, 255 registers, state / integrals / "
                   "generator in registers, 34 shuffles per mat-vec, shared memory unused in the sweep)" if k2_name != "misti_jsfs_kernel"
                   else "shared-memory / shuffle pipe (ncu: 81 % of peak), FP64 pipe 32 %")}
    dom = kernels["misti_correct_kernel" if dom_is_k1 else "misti_jsfs_kernel"]
    tr = ex.get("dram_bytes", {}).get("misti_correct_kernel" if dom_is_k1 else k2_name)
    traffic = None if not tr else tr["read"] + tr["write"]  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, one launch
    alg_bytes = B * (8 + 8 + 56 + 4)  # per item: parameter in, llh + spectrum + status out
    roofline = {"bound": "fp64", "kernel": dom["kernel"], "achieved": dom["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": dom["frac_of_fp64_peak"], "traffic": traffic,
                "note": "dominant kernel: FP64 operations it EXECUTED (thread-level 2*DFMA + DMUL + DADD from the SASS page of the "
                        "committed ncu capture, %s) x items / its CUDA-event time, against the measured DFMA peak; the kernel is not "
                        "at a throughput roof: see `kernels[...].bound` (what ncu shows)" % src,
                "kernels": kernels,
                "pair": {"ms": k1 + k2, "tflops": None if None in (k1_flops, k2_flops) else (k1_flops + k2_flops) * B / ((k1 + k2) * 1e-3) / 1e12,
                         "frac_of_fp64_peak": None if None in (k1_flops, k2_flops) else (k1_flops + k2_flops) * B / ((k1 + k2) * 1e-3) / 1e12 / peak},
                "dense_equivalent": {"flops_per_eval": F_DENSE, "tflops": F_DENSE * B / ((k1 + k2) * 1e-3) / 1e12,
                                     "frac_of_fp64_peak": F_DENSE * B / ((k1 + k2) * 1e-3) / 1e12 / peak,
                                     "note": "SURVEY 8d's dense formulation (Pade-13 + Van Loan on 45x45 / 9x9) over the kernel pair: not a "
                                             "utilisation -- the path runs closed forms and a sparse uniformisation, ~440x fewer FLOPs"},
                "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / ((k1 + k2) * 1e-3) / 1e9,
                        "peak_gbs": _measured_hbm(), "note": "HBM is not a bound of this path"},
                "terms_per_eval": terms_mean,
                "peak_source": "profiles/r01_fp64_peak.json (tools/fp64_peak.cu on this pool's B200: DFMA %.1f, DMMA %.1f, cuBLAS DGEMM "
                               "%.1f TFLOP/s; MEASURED_PEAKS.json has no FP64 entry)" % (peaks["dfma_tflops"], peaks["dmma_tflops"],
                                                                                          peaks["cublas_dgemm_tflops"])}
    cpu = None
    if world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        kind = cpu_kind()
        n = (6 if kind == "reference" else 128) * cores
        rate, dt = cpu_rate(n, cores)
        cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": kind,
               "sample": "%d of the batch's parameter vectors through %s, %d worker processes, %.1f s" % (n, cpu_path_text(), cores, dt)}
    d2h = world * B * 8 + B * 56 + B * 4
    line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": bench_config(B), "ok_fraction": ok_frac,
            "value_no_collective": world * B * args.steps / (nocoll_ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "returns": "llh of every item of the job (all-gathered at N > 1), expected JSFS [7] and status of this rank's items"},
            "gpu_launches": launches, "roofline": roofline, "batch_sweep": sweep, "config_sweep": cfgs, "likelihood_stage": lstage, "cpu_baseline": cpu, "time_to_fit": ttf, "clocks": clocks}
    _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _measured_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs")
    except (OSError, ValueError):
        return 6650.0  # B200_PROFILING.md fallback


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="parameter vectors per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the CPU legs (profiling runs)")
    ap.add_argument("--skip-fits", action="store_true", help="omit the time-to-fit legs (profiling runs)")
    ap.add_argument("--no-zero-copy", action="store_true", help="e2e at N = 1 through staged copies instead of pinned buffers the kernels access directly")
    ap.add_argument("--walkers", type=int, default=1024, help="basin-hopping walkers per GPU (config 3)")
    ap.add_argument("--fit-niter", type=int, default=100, help="basin-hopping hops per walker (config 3; the reference: 100)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # Exactly ONE line goes to stdout: while the benchmark runs, file descriptor 1 points to stderr (NCCL prints its
    # version banner to stdout from C, whatever NCCL_DEBUG_FILE says); the JSON line is written to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    _emit = lambda text: os.write(real_stdout, (text + "\n").encode())  # noqa: E731
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)


if __name__ == "__main__":
    main()
