#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: the per-direction diffuse-radiation sweep.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run)
    python bench.py --impl reference ...                      (CPU arm: the reference algorithm on the host cores)

One *step* = one full diffuse solve (all 12*4**(nAngularLevel-1) = 192 directions) over one synthetic grid:
computeOpacities + sweep + merge [+ NCCL all-reduce of Jmean1..3 when N > 1].  Default workload: the configuration the
metric's target is quoted on -- a 256^3 uniform grid, 192 directions (BASELINE.json configs[3], one sweep of it).
Directions are sharded across ranks (every GPU holds the whole grid), so per-GPU work shrinks with N: "strong".

Prints ONE JSON line (rank 0).  `value` = ray-cell segment updates per second with inputs resident in HBM;
`e2e` = the same metric through the host-buffer API (H2D of HI/HeI/HeII from pinned memory, D2H of Jmean1..3) inside the
timed region; `roofline` = algorithmic bytes (72 B per leaf per direction) of the sweep kernel launches
over their CUDA-event time, against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle (a port of the
reference: the Fortran itself cannot be built here) on a bounded sample of the same workload, one core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ray-cell segment updates/sec (diffuse sweep)"
METRIC_POINT = "ray-cell segment updates/sec (point-source ray casting + rate deposition)"
UNIT = "segment-updates/s"

# DRAM traffic of the dominant kernel, from one `ncu --set full` capture (profiles/): dram__bytes_read.sum +
# dram__bytes_write.sum of ONE launch, bytes (compare with roofline.algorithmic_bytes_per_launch)
NCU_TRAFFIC_BYTES = {
    "diffuse-256^3-uniform-192dir": 324.15e6 + 312.43e6,   # sweep_cell_kernel, one layer of all 32 zone tasks (r01e)
    "point-128^3-amr-100src": 5.51e6 + 2.66e6,             # point_march_kernel, pixel level 6 of 100 sources (r01b)
    "diffuse-128^3-amr2-192dir": 95.38e6 + 27.42e6,        # amr_wave_kernel, one mid-sweep wave of 878 (r01d); waves differ in size
}

WORKLOADS = {
    "diffuse-256^3-uniform-192dir": 256,
    "diffuse-128^3-uniform-192dir": 128,
    "diffuse-64^3-uniform-192dir": 64,
    # config 4 style: the reference's outer loop, sweep -> ionisation equilibrium, 10 passes per step, all on the device
    "iterate10-256^3-uniform-192dir": ("iterate", 256, 10),
    "iterate10-64^3-uniform-192dir": ("iterate", 64, 10),
    "iterate10-64^3-amr3-192dir": ("iterate", 64, 10, 3),     # the same loop on the config-5 style nested grid
    # config-5 style nested grid: n^3 base + refinement levels around a synthetic disc (general octree path)
    "diffuse-128^3-amr2-192dir": ("amr", 128, 2),
    "diffuse-64^3-amr3-192dir": ("amr", 64, 3),
    # point sources: n^3 base grid + one refined level over the central (n/4)^3 base cells, sources inside it
    "point-128^3-amr-100src": (128, 100),
    "point-256^3-amr-1000src": (256, 1000),
    "point-32^3-uniform-1src": (32, 1),
}


def make_inputs(spec, seed=1):
    """(n, grid dict, background) of a diffuse workload: uniform n^3 or a nested grid"""
    from radiativetransfer_b200 import workloads as W
    if isinstance(spec, tuple) and spec[0] == "iterate":
        if len(spec) > 3:   # nested grid: spec[3] refinement levels around the synthetic disc
            return spec[1], W.nested_grid(spec[1], spec[3], W.disc_refine(spec[3]), seed=5), W.uvb_background(3.0)
        return spec[1], W.uniform_grid(spec[1], seed=seed), W.uvb_background(3.0)
    if isinstance(spec, tuple):
        _, n, levels = spec
        return n, W.nested_grid(n, levels, W.disc_refine(levels), seed=5), W.uvb_background(3.0)
    return spec, W.uniform_grid(spec, seed=seed), W.uvb_background(3.0)


# ------------------------------------------------------------------------------------------------------------------
CLOCK_PERIOD_S = float(os.environ.get("RTB_BENCH_CLOCK_PERIOD", "0.1"))   # seconds between NVML clock samples


def sample_clocks(stop, out, dev):
    """SM clock and throttle reasons of GPU `dev` every CLOCK_PERIOD_S until `stop` is set.  NVML in-process (nvidia_ml_py): a
    query costs microseconds and does not fork -- spawning nvidia-smi from a process with a CUDA context stalls the
    launching thread for milliseconds, which is visible in short timed regions.  nvidia-smi is the fallback."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(int(dev))
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        bits = [0x8, 0x40, 0x20, 0x4]          # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while True:
            r = int(reasons_fn(h))
            out.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] +
                       ["Active" if r & b else "Not Active" for b in bits])
            if stop.wait(CLOCK_PERIOD_S):
                break
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(dev)],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
            "samples": len(samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------
class CpuDiffuse:
    """CPU oracle of the diffuse sweep on the workload's grid; the octree (and the worker threads' private copies) is
    built once, every sample() sweeps a bounded number of directions"""

    def __init__(self, n, grid, bg, threads):
        from oracle import ftte_oracle as fo
        self.og = fo.OracleGrid(n, grid["level"], grid["HI"], grid["HeI"], grid["HeII"], box_size=grid["box_size"])
        self.bg, self.n, self.threads = bg, n, threads
        self.order = (np.arange(192) * 37) % 192          # spread the sampled directions over the zones
        self.cursor = 0
        if threads > 1:  # creates the worker threads' private octree copies (set-up, untimed)
            self.og.diffuse_mt(bg["uvb"], bg["beta"], self.order[:0], nthreads=threads)

    def sample(self, seconds, per_batch=None):
        """sweeps batches of `threads` directions until `seconds` are used up (at least one batch)"""
        t0 = time.perf_counter()
        nseg, ndirs = 0, 0
        while True:
            rays = self.order[(self.cursor + np.arange(self.threads)) % 192]
            self.cursor = (self.cursor + self.threads) % 192
            o = self.og.diffuse_mt(self.bg["uvb"], self.bg["beta"], rays, nthreads=self.threads)
            assert o["status"] == 0
            nseg += o["nseg"]; ndirs += rays.size
            dt = time.perf_counter() - t0
            if dt + dt / (ndirs / self.threads) > seconds:
                break
        return nseg / dt, (f"{ndirs} direction sweep(s) out of the workload's 192 (cycled), same {self.n}^3 grid, "
                           f"{dt:.1f} s, {self.threads} thread(s)")


def cpu_oracle_sample(n, grid, bg, seconds, threads):
    """times the CPU oracle on a bounded number of directions of the same grid; returns (updates/s, description)"""
    return CpuDiffuse(n, grid, bg, threads).sample(seconds)


def point_inputs(workload):
    from radiativetransfer_b200 import workloads as W
    n, nsrc = WORKLOADS[workload]
    g, src = W.point_workload(n, nsrc, uniform="uniform" in workload)
    return n, g, src, np.ones(src.size, dtype=np.int32), W.synthetic_spectra()


def cpu_point_sample(g, src, wt, sp, seconds, threads):
    """times the CPU oracle (libm, as the reference) on a bounded number of the workload's sources, one source per
    host thread at a time (each thread a private copy of the octree)"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ftte_oracle as fo
    n = g["nx"]
    grids = [fo.OracleGrid(n, g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
             for _ in range(threads)]

    def one(i):
        return grids[i % threads].point(sp, src[i:i + 1], wt[i:i + 1])["nseg"]

    t0 = time.perf_counter()
    done, nseg = 0, 0
    with ThreadPoolExecutor(threads) as ex:    # ctypes releases the GIL inside the oracle call
        while done < src.size and (done == 0 or time.perf_counter() - t0 < seconds):
            chunk = list(range(done, min(done + threads, src.size)))
            nseg += sum(ex.map(one, chunk))
            done += len(chunk)
    dt = time.perf_counter() - t0
    return nseg / dt, f"{done} of {src.size} sources of the same grid, {dt:.1f} s, {threads} thread(s)"


def run_reference(args):
    """CPU arm: the reference algorithm (C++ oracle port; the Fortran cannot be compiled in this image) on the host
    cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload.startswith("point"):
        n, g, src, wt, sp = point_inputs(args.workload)
        threads = args.cpu_threads or int(max(1, min(os.cpu_count() or 1, 32)))
        vals = []
        for it in range(args.warmup + args.steps):
            v, desc = cpu_point_sample(g, src, wt, sp, max(2.0, min(20.0, 150.0 / (args.warmup + args.steps))), threads)
            if it >= args.warmup:
                vals.append(v)
        value = float(np.mean(vals))
        print(json.dumps({
            "impl": "reference", "metric": METRIC_POINT, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "leaves": int(g["level"].size), "sources": int(src.size),
                       "note": "each step = a bounded sample of sources of the workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    n, grid, bg = make_inputs(WORKLOADS[args.workload])
    import psutil
    cores = os.cpu_count() or 1
    avail = psutil.virtual_memory().available
    per_copy = 200.0 * int(grid["level"].size) * 1.3          # bytes per private octree copy
    # all host cores (every thread sweeps its own directions on a private copy of the octree); measured on the 16-core
    # GPU box at 256^3: 4 threads 1.3e7, 8 threads 2.1e7, 16 threads 3.0e7 segment updates/s
    threads = int(max(1, min(cores, 32, (0.5 * avail) // per_copy)))
    if args.cpu_threads:
        threads = args.cpu_threads
    cpu = CpuDiffuse(n, grid, bg, threads)
    vals = []
    desc = ""
    budget = max(2.0, min(20.0, 120.0 / (args.warmup + args.steps)))
    for it in range(args.warmup + args.steps):
        v, desc = cpu.sample(budget)
        if it >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    nseg_full = None
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "grid": f"{n}^3 base, {int(grid['level'].size)} leaves", "directions": 192,
                   "note": "each step = a bounded sample of directions of the workload, scaled per segment update"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def run_point(args):
    """point-source workloads: sources are sharded across the ranks (every GPU holds the whole grid), the six per-leaf
    rate fields are summed with one all-reduce"""
    import torch
    import torch.distributed as dist

    import radiativetransfer_b200 as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the rtb200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    n, g, src, wt, sp = point_inputs(args.workload)
    N = int(g["level"].size)
    eng = rt.Transport(device=local)
    eng.set_grid(n, g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    mine = slice(rank, None, world)                       # round-robin over the (sorted) source list
    R = torch.zeros(6, N, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        R.zero_()                                          # setZeroRates (equiSources.f90:1246)
        nseg = eng.point_device(sp, src[mine], wt[mine], R.data_ptr(), stream=stream)
        if world > 1:
            dist.all_reduce(R)
        return nseg

    for _ in range(W):
        nseg_rank = step()
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples, local), daemon=True)
    if rank == 0:
        th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        nseg_rank = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = eng.last_stats()
    barrier()
    t = torch.tensor([ms, float(nseg_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, nseg_total = float(tmax[0]), float(tsum[1])
    else:
        nseg_total = float(nseg_rank)
    ms_per_step = ms / args.steps
    value = nseg_total / (ms_per_step * 1e-3)

    # end to end through the host-buffer C-ABI call: H2D of the six rate arrays, D2H of them and of the diagnostics
    hRt = torch.zeros(6, N, dtype=torch.float64).pin_memory()    # the caller's rate fields, pinned host memory
    hR = hRt.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    eng.point(sp, src[mine], wt[mine], rates=hR, inplace=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hR[:] = 0.0                                              # setZeroRates on the host copy
        out = eng.point(sp, src[mine], wt[mine], rates=hR, inplace=True)
        if world > 1:
            Rt = torch.from_numpy(out["rates"]).to(dev)
            dist.all_reduce(Rt)
            out["rates"] = Rt.cpu().numpy()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt[0])
    stop.set()
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = st["algorithmic_bytes"] / (st["device_ms"] * 1e-3) / 1e9
        line = {
            "metric": METRIC_POINT, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "leaves": N, "sources": int(src.size), "max_pixel_level": 6,
                       "dust_approximation": 0, "segment_updates_per_step": nseg_total, "math": "fast",
                       "parallelism": f"sources sharded over {world} GPU(s), full grid per GPU, all-reduce of 6 rate fields",
                       "l2_policy": "gather workload: grid arrays larger than L2 at 256^3, tables L2-resident"},
            "clocks": clocks_summary(samples),
            "e2e": {"value": nseg_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 6 * N * 8, "d2h_bytes_per_step": 6 * N * 8 + int(src[mine].size) * 315 * 8},
            "gpu_launches": int(st["launches"]) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.workload) if world == 1 else None,
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "peak_source": peak_src, "kernel": "rtb::point_march_kernel",
                         "kernel_ms_per_step": st["device_ms"],
                         "note": "136 B per segment (5 reads + 6 read-modify-writes, SURVEY.md 8d); the path is bound "
                                 "by fp64 issue and gather latency, not by HBM (profiles/)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, desc = cpu_point_sample(g, src, wt, sp, args.cpu_seconds, 1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rtb200", choices=["rtb200", "reference"])
    ap.add_argument("--workload", default="diffuse-256^3-uniform-192dir", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload.startswith("point"):
        run_point(args)
        return

    import torch
    import torch.distributed as dist

    import radiativetransfer_b200 as rt
    from radiativetransfer_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the rtb200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)

    n, grid, bg = make_inputs(WORKLOADS[args.workload])
    N = int(grid["level"].size)
    spec = WORKLOADS[args.workload]
    iterations = spec[2] if isinstance(spec, tuple) and spec[0] == "iterate" else 0
    uniform = not isinstance(spec, tuple) or (iterations > 0 and len(spec) == 3)
    eng = rt.Transport(device=local)
    eng.set_grid(n, grid["level"], grid["HI"], grid["HeI"], grid["HeII"], grid["rho"], grid["abun2"], grid["box_size"])
    shards = sharding.shard_directions(world, n_angular_level=3, nx=n)
    rays = shards[rank] if world > 1 else None
    J = torch.zeros(3, N, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = torch.zeros(3, N, dtype=torch.float64, device=dev)     # krate24, krate25, krate26

    ksi_all = np.concatenate([bg["ksi24"], bg["ksi25"], bg["ksi26"]])
    if iterations:
        from radiativetransfer_b200 import workloads as Wk
        eng.set_rate_tables(Wk.rate_tables(5000))
        eng.set_temperature(10.0 ** np.random.default_rng(2).uniform(3.8, 4.6, N))

    def step_resident():
        if iterations:
            # the reference's outer loop (equiSources.f90:1230-1843) without its I/O: sweep -> solveRateEquations,
            # `iterations` passes; HI, HeI, HeII, J never leave the device.  Every step starts from the same state.
            eng.update_species(grid["HI"], grid["HeI"], grid["HeII"])
            total = 0
            for _ in range(iterations):
                total += eng.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), rays=rays, stream=stream)
                if world > 1:
                    dist.all_reduce(J)
                eng.chemistry_device(0, J.data_ptr(), ksi=ksi_all, stream=stream, want_change=False)
            return total
        nseg = eng.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), rays=rays, stream=stream)
        if world > 1:
            ar0.record()
            dist.all_reduce(J)            # per-leaf Jmean1..3 summed over the ranks' direction shards (NCCL, NVLink)
            ar1.record()
        # the diffuse contribution to the photo-rates (equiSources.f90:3546-3553) from the summed J
        eng.diffuse_rates_device(J.data_ptr(), bg["ksi24"], bg["ksi25"], bg["ksi26"], K[0].data_ptr(), K[1].data_ptr(),
                                 K[2].data_ptr(), stream=stream)
        return nseg

    # ---- resident-data timing: W warm-ups, then exactly K steps between barrier + synchronize ----
    for _ in range(W):
        nseg_rank = step_resident()
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples, local), daemon=True)
    if rank == 0:
        th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sweep_ms, sweep_launches, launches = 0.0, 0, 0
    e0.record()
    for _ in range(args.steps):
        nseg_rank = step_resident()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the library's own CUDA events (same stream) bracket the sweep-kernel launches of the LAST step
    st = eng.last_stats()
    sweep_ms, sweep_launches, launches = st["sweep_ms"], st["sweep_launches"], st["launches"]
    if sweep_ms <= 0:                       # general octree path: the library times the whole call
        sweep_ms = st["device_ms"]
    alg_bytes_rank = st["algorithmic_bytes"]
    barrier()
    t = torch.tensor([ms, float(nseg_rank), sweep_ms, alg_bytes_rank], dtype=torch.float64, device=dev)
    rank_kernel_ms = [sweep_ms]
    allreduce_ms = ar0.elapsed_time(ar1) if (world > 1 and not iterations) else 0.0   # last step; incl. waiting for the slowest rank
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, nseg_total = float(tmax[0]), float(tsum[1])
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_kernel_ms = [float(x[2]) for x in allt]
    else:
        nseg_total = float(nseg_rank)
    ms_per_step = ms / args.steps
    value = nseg_total / (ms_per_step * 1e-3)

    # ---- end-to-end through the host-buffer API: H2D of the species arrays + D2H of J inside the timed region ----
    hHI = torch.from_numpy(grid["HI"]).pin_memory()
    hHeI = torch.from_numpy(grid["HeI"]).pin_memory()
    hHeII = torch.from_numpy(grid["HeII"]).pin_memory()
    hJ = torch.empty(3, N, dtype=torch.float64).pin_memory()
    Jd = torch.zeros(3, N, dtype=torch.float64, device=dev)

    def step_e2e():
        eng.update_species(hHI.numpy(), hHeI.numpy(), hHeII.numpy())          # H2D from pinned host memory
        if iterations:
            for _ in range(iterations):
                eng.diffuse_device(bg["uvb"], bg["beta"], Jd.data_ptr(), rays=rays, stream=stream)
                if world > 1:
                    dist.all_reduce(Jd)
                eng.chemistry_device(0, Jd.data_ptr(), ksi=ksi_all, stream=stream, want_change=False)
            return float(eng.get_species()[0][0])                             # D2H of the new HI, HeI, HeII
        if world > 1:
            eng.diffuse_device(bg["uvb"], bg["beta"], Jd.data_ptr(), rays=rays, stream=stream)
            dist.all_reduce(Jd)
            hJ.copy_(Jd, non_blocking=False)                                   # D2H of the reduced result
        else:
            eng.diffuse(bg["uvb"], bg["beta"], out=hJ.numpy())                 # host-pointer C-ABI call, D2H inside
        return float(hJ[0, 0])

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt[0])
    e2e_value = nseg_total / (e2e_ms * 1e-3)
    stop.set()

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: sweep_cell_kernel; algorithmic bytes of this rank's launches over their event time
        achieved = alg_bytes_rank / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload,
                       "grid": f"{n}^3 uniform, lognormal tau" if uniform else f"{n}^3 base + nested levels (disc), {N} leaves",
                       "directions": 192,
                       "n_angular_level": 3, "frequency_groups": 3, "leaves": N,
                       "segment_updates_per_step": nseg_total, "math": "fast",
                       "step": (f"{iterations} x (computeOpacities + 192-direction sweep + merge [+ all-reduce of J] + "
                                "solveRateEquations), species re-uploaded at the start of the step") if iterations else
                               "computeOpacities + 192-direction sweep + merge [+ all-reduce of J] + diffuse photo-rates",
                       "parallelism": f"directions sharded over {world} GPU(s), full grid per GPU, all-reduce of J",
                       "l2_policy": "inputs larger than L2 (kappa + J + planes >> 126 MB)" if n >= 200 else
                                    "working set comparable to L2; not flushed between steps"},
            "clocks": clocks_summary(samples),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 3 * N * 8,
                    "d2h_bytes_per_step": 3 * N * 8},
            "gpu_launches": (int(launches) + 1) * max(iterations, 1) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.workload) if world == 1 else None,
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "algorithmic_bytes_per_launch": alg_bytes_rank / max(1, int(sweep_launches) - 1),
                         "peak_source": peak_src,
                         "kernel": "rtb::sweep_cell_kernel" if uniform else "rtb::amr_wave_kernel",
                         "launches_per_step": int(sweep_launches),
                         "algorithmic_bytes_per_step_this_rank": alg_bytes_rank, "kernel_ms_per_step": sweep_ms,
                         "kernel_ms_per_step_all_ranks": rank_kernel_ms,
                         "allreduce_ms_rank0_last_step": allreduce_ms,
                         "allreduce_bytes": 3 * N * 8 if world > 1 else 0,
                         "note": "72 B per leaf per direction; zones are swept with their directions fused, so DRAM "
                                 "traffic differs from the algorithmic bytes (see profiles/)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, desc = cpu_oracle_sample(n, grid, bg, args.cpu_seconds, 1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}
            except Exception as e:  # the checker is optional for the measurement itself
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
