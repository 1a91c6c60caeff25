#!/usr/bin/env python3
"""bench.py -- batched expected-JSFS + composite-logL evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of B synthetic parameter vectors per GPU
(BASELINE config 2: split index 40, unfolded SFS, one optimised migration band `-mi 2 5 12 0.8 1`,
`--cpfit`, m ~ U(0,5), numpy default_rng(1234 + rank)).  Prints ONE JSON line.

  value      evaluations/s, whole job, inputs resident in HBM, CUDA-event timed per step
  e2e        the same metric through the public host-buffer API (Engine.evaluate): pinned host
             buffers, H2D of the parameters and D2H of llh + status inside the timed region
  roofline   the dominant kernel against the measured FP64 peak (dense-equivalent algorithmic FLOPs of
             SURVEY.md 8d AND the FLOPs actually executed, counted from the kernel's own term counter)
  time_to_fit   one Nelder-Mead fit of the bench model stepped on the device (misti_nelder_mead) next to the same fit by
             the CPU oracle on one core (the second half of the BASELINE metric; rank 0, N = 1 only)
  cpu_baseline  the CPU oracle (oracle/misti_oracle.py, numpy/scipy port of the reference path) timed on
             the host cores on a bounded sample of the same parameter vectors (rank 0, N = 1 only)

--impl reference times the CPU implementation of the same path on the host cores (the reference is
pure Python and cannot travel to the GPU box; the oracle port is the same algorithm on the same scipy
calls), with one worker process per core.
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
# executed FLOPs per sparse mat-vec term of the uniformisation kernel (misti_jsfs.cuh inner loop):
# 44 rows x (1 diagonal + 4 off-diagonal slots) FMAs + 2 FMAs (P1, integral) + 1 per row
F_TERM = 44 * (2 * 5 + 2 * 2 + 1)
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
EXEC_FLOPS_FILE = os.path.join(ROOT, "profiles", "r01_executed_flops.json")
DRAM_TRAFFIC = None  # bytes per launch pair from the committed ncu --set full capture (set below when profiles/ has it)
if os.path.exists(EXEC_FLOPS_FILE):
    try:
        with open(EXEC_FLOPS_FILE) as _f:
            DRAM_TRAFFIC = json.load(_f).get("dram_bytes_per_launch_pair")
    except (OSError, ValueError):
        DRAM_TRAFFIC = None


def load_dataset():
    with open(os.path.join(ROOT, "tests", "golden", "datasets.json")) as f:
        return json.load(f)["datasets"]["synthetic"]


def make_params(B, rank):
    import numpy as np
    return np.random.default_rng(1234 + rank).uniform(0.0, 5.0, (B, 1))


def workload_name(B):
    return ("config2: m1.psmc m2.psmc m.sfs st=40 -uf -mi 2 5 12 0.8 1 --cpfit (numT=127), batched objective, "
            "B=%d vectors/GPU, m~U(0,5)" % B)


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port): one worker process per core
# ------------------------------------------------------------------------------------------------
_worker_model = None


def _worker_eval(m):
    global _worker_model
    if _worker_model is None:
        from oracle.misti_oracle import OracleModel
        ds = load_dataset()
        _worker_model = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], SPLIT_T, [BAND], [], cpfit=True, smooth=True,
                                    unfolded=True)
    return float(_worker_model.likelihood([m]))


def cpu_rate(n_evals, cores, rank=0):
    """evals/s of the CPU oracle on `cores` worker processes over n_evals parameter vectors."""
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
    per_step = 16 * cores
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
    import numpy, scipy
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "cpu_path": "oracle/misti_oracle.py (numpy %s / scipy %s port of "
                       "the reference path; the reference is pure Python and is absent on the GPU box)" % (numpy.__version__, scipy.__version__)},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
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
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.monotonic()] + [c.strip() for c in line.split(",")])

    def stop(self, t0, t1):
        """median SM clock over the samples taken inside [t0, t1] (the timed regions)"""
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        rows = [r[1:] for r in self.rows if t0 <= r[0] <= t1 + 0.05]
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


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import misti_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not land on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(dev)  # a real (non-NULL) stream shared by torch events and the engine's launches
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

    def step_device():
        eng.evaluate_device(B, 1, params_d.data_ptr(), llh_d.data_ptr(), model=mid, flags=flags, jafs_ptr=jafs_d.data_ptr(),
                            status_ptr=status_d.data_ptr(), terms_ptr=terms_d.data_ptr())

    # ---- device-resident throughput -----------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    k1_ms, k2_ms = [], []
    barrier()
    t_load0 = time.monotonic()
    for s in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
        ev[s][0].record(stream)
        step_device()
        ev[s][1].record(stream)
        a, b = eng.last_kernel_ms()  # synchronises with this step
        k1_ms.append(a)
        k2_ms.append(b)
    barrier()
    launches = eng.launch_count() - launches0
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)

    # ---- end to end through the public host-buffer API -----------------------------------------
    llh_h = torch.empty((B, 1), dtype=torch.float64).pin_memory()
    status_h = torch.empty((B,), dtype=torch.int32).pin_memory()
    bufs = {"llh": llh_h.numpy(), "status": status_h.numpy()}
    p_np = params_h.numpy()
    for _ in range(max(1, args.warmup)):
        eng.evaluate(p_np, model=mid, flags=flags, want=("status",), buffers=bufs)
    barrier()
    ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.zero_()
        ee[s][0].record(stream)
        eng.evaluate(p_np, model=mid, flags=flags, want=("status",), buffers=bufs)
        ee[s][1].record(stream)
    barrier()
    e2e_ms = sum(e0.elapsed_time(e1) for e0, e1 in ee)
    clocks = sampler.stop(t_load0, time.monotonic()) if rank == 0 else None

    ok_frac = float((status_d == 0).float().mean().item())
    terms_mean = float(terms_d.double().mean().item())
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only data-path collective of the workload: gather the small likelihood vectors to every rank
        gathered = torch.empty((world * B, 1), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, llh_d)
    total_ms, e2e_ms = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    k1, k2 = sum(k1_ms) / len(k1_ms), sum(k2_ms) / len(k2_ms)
    dom = "misti_jsfs_kernel" if k2 >= k1 else "misti_correct_kernel"
    with open(FP64_PEAK_FILE) as f:
        peaks = json.load(f)
    peak = peaks["dfma_tflops"]  # the executed kernels are DFMA code
    # One evaluation = one item through BOTH kernels (correction chain, then JSFS + likelihood); the algorithmic figure of
    # SURVEY 8d belongs to the evaluation, so it is set against the device time of the pair.
    pair_ms = k1 + k2
    dense_tflops = F_DENSE * B / (pair_ms * 1e-3) / 1e12
    k2_flops = F_TERM * terms_mean  # counted by the kernel: mat-vec terms (a zero-migration run counts as one)
    k1_flops, k1_src = None, None
    if os.path.exists(EXEC_FLOPS_FILE):
        with open(EXEC_FLOPS_FILE) as f:
            ex = json.load(f)
        k1_flops = ex["misti_correct_kernel"]["flops_per_item"]
        k1_src = "profiles/r01_executed_flops.json (ncu SASS instruction counts of this workload: thread-level 2*DFMA + DMUL + DADD)"
    exec_flops = k2_flops + (k1_flops or 0.0)
    exec_tflops = exec_flops * B / (pair_ms * 1e-3) / 1e12
    roofline = {"bound": "fp64", "kernel": "misti_correct_kernel + misti_jsfs_kernel (one evaluation = both; longer one: %s)" % dom,
                "achieved": dense_tflops, "peak": peak, "unit": "TFLOP/s", "frac": dense_tflops / peak, "traffic": DRAM_TRAFFIC,
                "note": "achieved = dense-equivalent algorithmic FLOPs of SURVEY 8d (%.1f MFLOP/eval: Pade-13 + Van Loan on 45x45 / "
                        "9x9) / device time of the kernel pair; the path instead runs closed forms for zero-migration runs and the "
                        "post-split tail and a sparse uniformisation for the intervals with migration, so frac >> 1 is expected; "
                        "'executed' is what the FP64 pipe really did; traffic = dram bytes read + written per launch pair "
                        "(ncu --set full, profiles/)" % (F_DENSE / 1e6),
                "executed": {"tflops": exec_tflops, "frac": exec_tflops / peak, "flops_per_eval": exec_flops,
                             "misti_jsfs_kernel": {"flops_per_eval": k2_flops, "terms_per_eval": terms_mean,
                                                   "tflops": k2_flops * B / (k2 * 1e-3) / 1e12,
                                                   "frac": k2_flops * B / (k2 * 1e-3) / 1e12 / peak},
                             "misti_correct_kernel": None if k1_flops is None else {
                                 "flops_per_eval": k1_flops, "tflops": k1_flops * B / (k1 * 1e-3) / 1e12,
                                 "frac": k1_flops * B / (k1 * 1e-3) / 1e12 / peak, "source": k1_src}},
                "kernel_ms": {"misti_correct_kernel": k1, "misti_jsfs_kernel": k2},
                "peak_source": "profiles/r01_fp64_peak.json (tools/fp64_peak.cu on this pool's B200: DFMA %.1f, DMMA %.1f, cuBLAS DGEMM "
                               "%.1f TFLOP/s; MEASURED_PEAKS.json has no FP64 entry)" % (peaks["dfma_tflops"], peaks["dmma_tflops"],
                                                                                          peaks["cublas_dgemm_tflops"])}
    cpu = None
    if world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        n = 128 * cores
        rate, dt = cpu_rate(n, cores)
        cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": "%d of the batch's parameter vectors through oracle/misti_oracle.py (numpy/scipy port of the reference "
                         "path), %d worker processes, %.1f s" % (n, cores, dt)}
    # second half of the BASELINE metric: time to fit.  One Nelder-Mead fit of the bench model (MiSTI.py ... -mi 2 5 12 0.8 1
    # --cpfit, tol 1e-4) stepped on the device, next to the same fit by the CPU oracle (scipy Nelder-Mead, one core).
    ttf = None
    if world == 1 and not args.skip_cpu:
        import numpy as np
        x0, one, zero = np.array([[BAND[3]]]), np.array([mid], dtype=np.int32), np.zeros(1, dtype=np.int32)
        eng.nelder_mead(x0, one, zero, flags=flags, xatol=1e-4, fatol=1e-4, maxiter=1000)
        t0 = time.perf_counter()
        fit = eng.nelder_mead(x0, one, zero, flags=flags, xatol=1e-4, fatol=1e-4, maxiter=1000)
        gpu_s = time.perf_counter() - t0
        from oracle.misti_oracle import OracleModel
        om = OracleModel(ds["times"], ds["lambdas"], ds["sfs"], SPLIT_T, [BAND], [], cpfit=True, smooth=True, unfolded=True)
        t0 = time.perf_counter()
        ref = om.solve(1e-4)
        cpu_s = time.perf_counter() - t0
        ref_x, ref_llh = ref[0], ref[1]
        ttf = {"config": "config2: one Nelder-Mead fit of the optimised band (tol 1e-4, start 0.8)", "gpu_s": gpu_s,
               "gpu_x": fit["x"][0].tolist(), "gpu_llh": float(-fit["fun"][0]), "gpu_nfev": int(fit["nfev"][0]),
               "gpu_rounds_of_launches": int(fit["launches"]), "cpu_s": cpu_s, "cpu_kind": "port", "cpu_cores": 1,
               "cpu_x": [float(v) for v in ref_x], "cpu_llh": float(ref_llh)}
    line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(B), "l2": "256 MiB buffer rewritten between timed iterations", "ok_fraction": ok_frac,
                       "parallelism": "independent items sharded across ranks; all_gather of llh only"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": B * 8 + B * 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "time_to_fit": ttf, "clocks": clocks}
    _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="parameter vectors per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
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
