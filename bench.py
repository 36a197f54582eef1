#!/usr/bin/env python
"""bench.py -- V-cycle throughput of the VariableCoeffPoissonOperator multigrid hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 512] [--smooth 2] [--box 32] [--keep-b] [--smoother 1]
  python bench.py --impl reference ...      # the CPU restatement (oracle) on the host cores, same metric

A "step" is one multigrid V-cycle on a zeroed correction (what [Chombo] MultilevelLinearOp::preCond does per
V-cycle): setToZero(e); MultiGrid::oneCycle(e, r) -- pre-smooth, restrictResidual, recurse, bottom BiCGStab,
prolongIncrement, post-smooth on every MG depth.  Synthetic data: the reference's params.txt Bowen-York binary
(Main_PoissonSolver.cpp first nonlinear iteration: psi = 1, rhs / aCoef from set_rhs / set_a_coef), single level.

Prints ONE JSON line (see the contract in the task description / DESIGN.md "Measurement").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vcycle_gdof_per_s_V22_512cubed_per_gpu"
UNIT = "GDOF/s"


def algorithmic_bytes_per_cell(smooth, keep_b):
    """per level 48(nu1+nu2)+52 B/cell with bCoef streamed; bCoef == 1 dropped: 40(nu1+nu2)+44."""
    sweep = 48 if keep_b else 40
    level = sweep * 2 * smooth + (33 + 1 + 1 + 17 if keep_b else 25 + 1 + 1 + 17)
    return sweep, level


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clocks / clock-event (throttle) reasons sampled DURING the timed region: an NVML thread (10 ms period; the
    timed region is ~100 ms) with the recipe's `nvidia-smi --query-gpu ... -lms` process beside it as the fallback.
    start() returns only once a first sample exists, so a slow nvidia-smi / NVML start-up cannot leave the window empty."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None          # nvidia-smi: (time, csv line)
        self.nv, self.nv_rows, self.nv_max, self.nv_stop = None, [], None, threading.Event()   # NVML: (time, MHz, mask)

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip()]
        if ids and self.index < len(ids) and ids[self.index].strip().isdigit():
            return int(ids[self.index])
        return self.index

    def start(self, wait_s=8.0):
        dev = self._nvml_index()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.nv_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nv_bits = (pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                            pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap)

            def loop():
                while not self.nv_stop.is_set():
                    try:
                        self.nv_rows.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(reasons(h))))
                    except Exception:
                        pass
                    self.nv_stop.wait(0.01)
            self.nv = threading.Thread(target=loop, daemon=True)
            self.nv.start()
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        deadline = time.time() + wait_s
        while time.time() < deadline and (self.nv or self.proc):
            if (self.nv is None or self.nv_rows) and (self.proc is None or self.rows):
                break
            if self.proc is not None and self.proc.poll() is not None and not self.rows:
                self.proc = None        # nvidia-smi exited without a row
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc and not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi and NVML unavailable"], "samples": 0}
        time.sleep(0.12)
        self.nv_stop.set()
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons, source = [], self.nv_max, set(), "nvml"
        for t, mhz, mask in self.nv_rows:
            if t0 <= t <= t1:
                sm.append(mhz)
                reasons.update(n for n, b in zip(self.NAMES, self.nv_bits) if mask & b)
        if not sm:                       # no NVML: the nvidia-smi rows (its timestamps lag the sample by up to one period)
            source = "nvidia-smi"
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                inside = t0 - 0.05 <= t <= t1 + 0.1
                try:
                    if inside:
                        sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                if inside:
                    reasons.update(n for n, val in zip(self.NAMES, f[4:8]) if val.lower().startswith("active"))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": source}


def host_mem_available_gib():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return None


def cpu_sample_n(args):
    """Side of the CPU problem: the workload's own n (512) -- the boxed oracle needs ~24 GiB of host memory for it
    (8 multigrid_vars components with three ghost layers per 32^3 box dominate); a smaller host gets n/2 and says so."""
    n = args.cpu_n if args.cpu_n else args.n
    avail = host_mem_available_gib()
    need = 24.0 * (n / 512.0) ** 3 + 6.0
    while avail is not None and avail < need and n > 64:
        n //= 2
        need = 24.0 * (n / 512.0) ** 3 + 6.0
    return n


def world_mult(world, scaling):
    """weak scaling: n^3 cells per GPU; the domain doubles in z, then y, then x:
    512^3 -> 512x512x1024 -> 512x1024x1024 -> 1024^3, dx constant (L is the x-length), z-slabs of nz/world planes.
    strong (config C3): the n^3 domain itself is cut into `world` z-slabs."""
    if scaling == "strong":
        return (1, 1, 1)
    return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world, (1, 1, world))


def workload_config(args, world):
    """A pure function of the command line and the GPU count: both arms print the SAME config."""
    mult = world_mult(world, args.scaling)
    N = [args.n * mult[0], args.n * mult[1], args.n * mult[2]]
    return {"workload": f"single-level Bowen-York binary (params.txt physics), {args.n}^3 cells per GPU "
                        f"({N[0]}x{N[1]}x{N[2]} on {world} GPU{'s' if world > 1 else ''}), V({args.smooth},{args.smooth}), "
                        f"max_grid_size {args.box}, harmonic coefficient averaging, Dirichlet dpsi=0",
            "n": args.n, "global_N": N, "global_cells": N[0] * N[1] * N[2], "numMGsmooth": args.smooth, "max_grid_size": args.box,
            "l2": "inputs larger than L2 (every level-0 array is n^3*8 B = %.0f MiB >> 126 MB); no flush needed" % (args.n ** 3 * 8 / 2 ** 20),
            "bCoef": "streamed" if args.keep_b else "dropped (identically 1; bit-identical results)",
            "smoother": "fused red+black sweep" if args.smoother == 1 else "one launch per colour",
            "decomposition": "z-slabs, one per GPU"}


def run_reference(args):
    """The reference's CPU path: the boxed C++ restatement under oracle/ (the reference itself needs Chombo 3.2 +
    Fortran + MPI and cannot be built here), OpenMP over boxes on ALL host cores (whatever OMP_NUM_THREADS says: torchrun
    exports 1), each step one V-cycle of the workload's per-GPU problem (n^3; GDOF/s is a throughput, so the n^3 sample
    stands for the N-GPU domain at N > 1)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import Oracle, use_all_host_cores
    cores = use_all_host_cores()
    n = cpu_sample_n(args)
    o = Oracle(N=(n, n, n), max_grid_size=args.box, numMGsmooth=args.smooth)
    o.setup()
    o.load_rhs_zero_e()
    zero = np.zeros((n, n, n))
    for _ in range(args.warmup):
        o.set("E", zero)
        o.vcycle()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.set("E", zero)     # setToZero(e), part of the step like on the GPU arm
        o.vcycle()
    dt = (time.perf_counter() - t0) / args.steps
    val = n ** 3 / dt / 1e9
    sample = (f"{args.steps} V({args.smooth},{args.smooth}) cycles at {n}^3 ({args.box}^3 boxes, boxed C++ oracle, OpenMP over boxes on "
              f"{cores} host threads)" + ("" if n == args.n else f"; the host has too little memory for {args.n}^3"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(world, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "n_timed": n},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Problem:
    """One Bowen-York level on this rank's z-slab: fields, the MG hierarchy (factory) and a correction."""

    def __init__(self, m, ctx, args, N, Lx, k0, nzl):
        self.m, self.ctx = m, ctx
        self.N, self.k0, self.nzl = N, k0, nzl
        self.P = m.make_params(dict(m.DEFAULTS, N=N, L=Lx, max_grid_size=args.box, numMGsmooth=args.smooth))
        self.lvl = m.level_op_from_params(ctx, self.P, k0, nzl)
        vars_ = m.MultigridVars(ctx, self.P, k0, nzl)
        self.dpsi, self.rhs, self.a, self.b = (self.lvl.create() for _ in range(4))
        vars_.set_initial_conditions(self.dpsi)
        vars_.set_rhs_and_a_coef(self.rhs, self.a)
        vars_.set_b_coef(self.b)
        vars_.close()
        self.f = m.VariableCoeffPoissonOperatorFactory(ctx, self.P, self.a, self.b, keep_b=args.keep_b)
        self.f.set_smoother(args.smoother)
        self.op = self.f.MGnewOp(0)
        self.e = self.op.create()
        self.extra = []

    def step(self):
        self.f.vcycle_from_zero(self.e, self.rhs)   # setToZero(e); MultiGrid::oneCycle(e, rhs)

    def close(self):
        self.ctx.sync()
        for x in self.extra + [self.e]:
            x.close()
        self.f.close()
        for x in (self.dpsi, self.rhs, self.a, self.b):
            x.close()
        self.lvl.close()


def cpu_baseline_and_parity(args, pb, ms_gpu_step):
    """N = 1, after the timed regions: the boxed oracle at the workload's own size on all host cores.  It is timed (the
    cpu_baseline record) AND used as the checker of the very launch shapes the bench timed (the parity record): with the
    oracle's coefficients uploaded, one finest-level fused sweep must equal the oracle's sweep bit for bit, and one
    V-cycle must agree to 1e-10 relative in max-norm with the same bottom BiCGStab iteration count."""
    from oracle import Oracle, use_all_host_cores
    cores = use_all_host_cores()
    n = cpu_sample_n(args)
    o = Oracle(N=(n, n, n), max_grid_size=args.box, numMGsmooth=args.smooth)
    o.setup()
    o.load_rhs_zero_e()
    it_o = o.vcycle()
    e_o = o.get("E")
    zero = np.zeros((n, n, n))
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        o.set("E", zero)
        o.vcycle()
    dt = (time.perf_counter() - t0) / reps
    base = {"value": n ** 3 / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "n_timed": n,
            "sample": f"{reps} V({args.smooth},{args.smooth}) cycles at {n}^3 ({args.box}^3 boxes) with the boxed C++ oracle, "
                      f"OpenMP over boxes on {cores} host threads", "ms_per_vcycle": dt * 1e3}
    parity = None
    if n == args.n and pb is not None:
        pb.a.upload(o.get("A"))
        pb.rhs.upload(o.get("RHS"))
        pb.f.refresh_coefs()
        pb.op.setToZero(pb.e)
        o.load_rhs_zero_e()
        pb.op.relax(pb.e, pb.rhs, 1)
        o.relax(0, 1)
        sweep_equal = bool(np.array_equal(pb.e.download(), o.get("E")))
        pb.f.vcycle_from_zero(pb.e, pb.rhs)
        it_g = pb.f.last_bottom_iterations
        e_g = pb.e.download()
        err = float(np.abs(e_g - e_o).max() / np.abs(e_o).max())
        pb.op.residual(pb.dpsi, pb.e, pb.rhs, True)
        o.set("E", e_o)
        res_o = float(np.abs(o.residual(0, True)).max())
        res_g = float(pb.op.norm(pb.dpsi, 0))
        parity = {"checker": f"CPU oracle (oracle/mgic_oracle.cpp) at the timed size {n}^3, same coefficients",
                  "finest_level_fused_sweep_bit_identical": sweep_equal,
                  "vcycle_rel_err_max_norm": err, "tolerance": 1e-10,
                  "bottom_iterations": [it_g, it_o], "residual_max_norm_after_cycle": [res_g, res_o],
                  "ok": bool(sweep_equal and err < 1e-10 and it_g == it_o and abs(res_g - res_o) <= 1e-10 * res_o + 1e-13 * float(np.abs(o.get("RHS")).max()))}
    o.close()
    return base, parity


def multirank_parity(m, ctx, args, dist, rank, world, local):
    """N > 1: a small problem of the same shape (128^3 cells per rank, same decomposition) is cycled twice on the N ranks
    and, independently, as ONE domain on every rank's own GPU (a second, single-rank context); the rank's slab of the
    correction must be bit-identical."""
    import torch
    ns = 128
    mult = world_mult(world, args.scaling if args.scaling == "weak" else "weak")
    N = (ns * mult[0], ns * mult[1], ns * mult[2])
    nzl = N[2] // world
    pm = Problem(m, ctx, args, N, 100.0 * mult[0], rank * nzl, nzl)
    for _ in range(2):
        pm.step()
    mine = pm.e.download()[rank * nzl:(rank + 1) * nzl].copy()   # the C ABI addresses global arrays: only the slab is filled
    its_m = pm.f.last_bottom_iterations
    pm.close()
    c1 = m.Context(local)
    p1 = Problem(m, c1, args, N, 100.0 * mult[0], 0, N[2])
    for _ in range(2):
        p1.step()
    whole = p1.e.download()
    its_1 = p1.f.last_bottom_iterations
    p1.close()
    c1.close()
    ok = bool(np.array_equal(mine, whole[rank * nzl:(rank + 1) * nzl])) and its_m == its_1 and bool(np.abs(mine).max() > 0)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"checker": f"the same {N[0]}x{N[1]}x{N[2]} V-cycles on ONE GPU (single-rank context), per rank",
            "slab_bit_identical_on_all_ranks": bool(t.item() == 1.0), "bottom_iterations": [its_m, its_1],
            "ok": bool(t.item() == 1.0)}


def kernel_fingerprint():
    """sha256 of the sources the dominant kernel is built from: the committed ncu traffic figure is only quoted for the
    build it was captured on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("gsrb_fused.cu", "mgic_device.cuh", "tma.cuh"):     # the kernel, the point update and the TMA wrappers it inlines
        h.update(open(os.path.join(ROOT, "mg_ic_code_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def profiled_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (average over the finest-level
    sweeps of one V-cycle, like `achieved`) from the committed `ncu --set full` capture of this command -- quoted only if
    that capture was taken on THIS build of the kernel (source fingerprint) and configuration; null otherwise."""
    p = os.path.join(ROOT, "profiles", TRAFFIC_SOURCE)
    if not (os.path.exists(p) and args.n == 512 and not args.keep_b and args.smoother == 1 and args.smooth == 2 and args.fused_cfg in (None, 4, 5)):
        return None
    try:
        d = json.load(open(p))
        if d.get("kernel_source_sha16") != kernel_fingerprint():
            return None
        return float(d["finest_level_avg_dram_gbyte_per_launch"]["total"]) * 1e9
    except Exception:
        return None


TRAFFIC_SOURCE = "r2_fused_vcycle_ncu_summary.json"


def time_steps(pb, steps, barrier, stream, torch):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record(stream)
    for _ in range(steps):
        pb.step()
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1), t0, time.time()


def run_gpu(args):
    import torch
    import mg_ic_code_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = m.Context(local, rank=rank, nranks=world)
    if world > 1:
        from mg_ic_code_b200 import comm
        comm.attach(ctx, dist)
    if args.halo == "nccl":
        ctx.set_option("p2p_halo", 0)
    if args.overlap_halo:
        ctx.set_option("overlap_halo", 1)
    if args.fused_cfg is not None:
        ctx.set_option("fused_cfg", args.fused_cfg)
    if args.fused_min_cells is not None:
        ctx.set_option("fused_min_cells", args.fused_min_cells)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))

    n = args.n
    mult = world_mult(world, args.scaling)
    N = (n * mult[0], n * mult[1], n * mult[2])
    nzl = N[2] // world
    k0 = rank * nzl
    pb = Problem(m, ctx, args, N, 100.0 * mult[0], k0, nzl)
    f, op, e, rhs = pb.f, pb.op, pb.e, pb.rhs
    cells_local = N[0] * N[1] * nzl
    cells_total = cells_local * world

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(max(args.warmup, 3)):
        pb.step()
    barrier()
    # ---- device-resident timing (value) ---------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = ctx.launch_count
    ms, t0, t1 = time_steps(pb, args.steps, barrier, stream, torch)
    launches = ctx.launch_count - l0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    bottom_iters = f.last_bottom_iterations
    # ---- same K steps again with a CUDA-event pair around every finest-level GSRB launch (the dominant kernel).
    # Event records cannot live inside a replayed CUDA graph, so this pass launches eagerly; its own total time is the
    # denominator of the kernel's share of the step.
    ctx.profile(True)
    ms_prof, _, _ = time_steps(pb, args.steps, barrier, stream, torch)
    k_launches, k_ms = ctx.profile_read()
    breakdown = {k: {"launches": v[0] // args.steps, "ms_per_step": v[1] / args.steps} for k, v in ctx.profile_breakdown().items()}
    ctx.profile(False)
    # ---- end-to-end through the C ABI with HOST buffers: H2D residual, V-cycle, D2H correction ----------------
    # each rank's pinned buffers hold its own slab; the C ABI addresses global arrays, so pass the slab-shifted base
    h_r = torch.empty(cells_local, dtype=torch.float64).pin_memory()
    h_e = torch.empty(cells_local, dtype=torch.float64).pin_memory()
    off = k0 * N[0] * N[1] * 8
    L = m.lib()
    m._capi.check(L.mgic_field_download_async(rhs.h, C.c_void_p(h_r.data_ptr() - off)))
    ctx.sync()
    # pipelined over two (residual, correction) buffer pairs: the upload of step i+1 runs on the H2D stream while step i
    # is computed and step i-1's correction goes back on the D2H stream (mgic_field_prefetch / _writeback / _wait)
    r_bufs, e_bufs = [op.create(), op.create()], [e, op.create()]
    pb.extra += r_bufs + e_bufs[1:]
    p_r, p_e = C.c_void_p(h_r.data_ptr() - off), C.c_void_p(h_e.data_ptr() - off)

    def run_e2e(k):
        m._capi.check(L.mgic_field_prefetch(r_bufs[0].h, p_r))
        for i in range(k):
            rb, eb = r_bufs[i % 2], e_bufs[i % 2]
            m._capi.check(L.mgic_field_wait(rb.h))          # its upload
            m._capi.check(L.mgic_field_wait(eb.h))          # the download of step i-2 out of this buffer
            if i + 1 < k:                                   # ordered after V-cycle i-1, the last reader of that buffer
                m._capi.check(L.mgic_field_prefetch(r_bufs[(i + 1) % 2].h, p_r))
            f.vcycle_from_zero(eb, rb)
            m._capi.check(L.mgic_field_writeback(eb.h, p_e))
        for eb in e_bufs:
            m._capi.check(L.mgic_field_wait(eb.h))

    def run_copies_only(k):
        """the same host<->HBM traffic with no V-cycle: what the PCIe / host-memory path alone sustains with both
        directions busy on every rank at once (the floor of the end-to-end step)"""
        for i in range(k):
            m._capi.check(L.mgic_field_prefetch(r_bufs[i % 2].h, p_r))
            m._capi.check(L.mgic_field_writeback(e_bufs[i % 2].h, p_e))
        for x in r_bufs + e_bufs:
            m._capi.check(L.mgic_field_wait(x.h))

    e2e_steps = max(4, min(args.steps, 10))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_e2e(2)                                              # captures the V-cycle graphs of both buffer pairs
    barrier()
    ev0.record(stream)
    run_e2e(e2e_steps)
    ev1.record(stream)
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    # the correction that came back is the one the device holds
    # (h_r is not needed any more: reuse it as the comparison buffer instead of pinning another GiB per rank)
    chk = torch.empty_like(h_e)
    m._capi.check(L.mgic_field_download_async(e_bufs[(e2e_steps - 1) % 2].h, C.c_void_p(chk.data_ptr() - off)))
    ctx.sync()
    if not torch.equal(chk, h_e) or not bool(torch.isfinite(h_e).all()) or float(h_e.abs().max()) == 0.0:
        raise SystemExit("bench.py: the end-to-end leg returned a wrong correction")
    del chk
    run_copies_only(2)
    barrier()
    tc0 = time.perf_counter()
    run_copies_only(e2e_steps)
    barrier()
    ms_copy = (time.perf_counter() - tc0) * 1e3
    # max over ranks
    if dist is not None:
        t = torch.tensor([ms, ms_e2e, k_ms, ms_prof, ms_copy], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, k_ms, ms_prof, ms_copy = t.tolist()
    halo_info = None
    if world > 1:
        from mg_ic_code_b200 import comm
        hs = comm.halo_stats(ctx)
        halo_info = {"transport": "NVLink peer stores (CUDA IPC)" if hs[0] > 0 and hs[1] == 0 else
                     ("ncclSend/ncclRecv" if hs[0] == 0 else "mixed"), "peer_store_exchanges": hs[0], "nccl_exchanges": hs[1],
                     "bytes_sent_rank0": comm.halo_bytes(ctx), "overlap_with_interior": bool(args.overlap_halo)}
    depths = f.depths
    del h_r, h_e
    # ---- N > 1: the other reading of BASELINE's metric, strong scaling of ONE n^3 domain (config C3), in the same run --
    strong = None
    parity = None
    if world > 1 and not args.no_strong and args.scaling == "weak":
        pb.close()
        pb = None
        nzs = n // world
        ps = Problem(m, ctx, args, (n, n, n), 100.0, rank * nzs, nzs)
        for _ in range(max(args.warmup, 3)):
            ps.step()
        ms_s, _, _ = time_steps(ps, args.steps, barrier, stream, torch)
        its_s = ps.f.last_bottom_iterations
        ps.close()
        # the one-GPU time of the same domain, measured now on rank 0's GPU with a single-rank context
        ms_1 = 0.0
        if rank == 0:
            c1 = m.Context(local)
            p1 = Problem(m, c1, args, (n, n, n), 100.0, 0, n)
            for _ in range(max(args.warmup, 3)):
                p1.step()
            c1.sync()
            st1 = torch.cuda.ExternalStream(c1.stream)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(st1)
            for _ in range(args.steps):
                p1.step()
            a1.record(st1)
            c1.sync()
            ms_1 = a0.elapsed_time(a1)
            p1.close()
            c1.close()
        t = torch.tensor([ms_s, ms_1], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_s, ms_1 = t.tolist()
        strong = {"what": f"config C3: ONE {n}^3 domain cut into {world} z-slabs, same V-cycle, same run",
                  "ms_per_step": ms_s / args.steps, "value": n ** 3 / (ms_s / args.steps * 1e-3) / 1e9, "unit": UNIT,
                  "one_gpu_ms_per_step_same_run": ms_1 / args.steps, "speedup_vs_one_gpu": ms_1 / ms_s,
                  "efficiency_vs_n1": ms_1 / ms_s / world, "bottom_bicgstab_iterations": its_s}
    if world > 1 and not args.no_parity:
        if pb is not None:
            pb.close()
            pb = None
        parity = multirank_parity(m, ctx, args, dist, rank, world, local)
    # ---- the same weak-scaling V-cycle with the reference's max_grid_size at 64 (params.txt key; it sets the MG depth,
    # Factory.cpp:168-172): one level more, so the coarsest -- at N > 1 agglomerated -- level is 8x smaller.  Reported
    # next to the headline configuration because that level is what costs the headline its parallel efficiency.
    alt = None
    if args.box != 64 and not args.no_alt and args.scaling == "weak":
        import types
        a2 = types.SimpleNamespace(**dict(vars(args), box=64))
        if world > 1 and pb is not None:
            pb.close()
            pb = None
        pa = Problem(m, ctx, a2, N, 100.0 * mult[0], k0, nzl)
        for _ in range(max(args.warmup, 3)):
            pa.step()
        ms_a, _, _ = time_steps(pa, args.steps, barrier, stream, torch)
        alt_iters, alt_depths = pa.f.last_bottom_iterations, pa.f.depths
        pa.close()
        if dist is not None:
            t = torch.tensor([ms_a], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_a = t.item()
        alt = {"what": "the same V-cycle with max_grid_size = 64 instead of %d (the reference's params.txt key; one MG level more)" % args.box,
               "ms_per_step": ms_a / args.steps, "value": cells_total / (ms_a / args.steps * 1e-3) / 1e9, "unit": UNIT,
               "mg_depths": alt_depths, "bottom_bicgstab_iterations": alt_iters}
    if rank == 0:
        ms_step = ms / args.steps
        value = cells_total / (ms_step * 1e-3) / 1e9
        e2e_val = cells_total / (ms_e2e / e2e_steps * 1e-3) / 1e9
        sweep_b, level_b = algorithmic_bytes_per_cell(args.smooth, args.keep_b)
        peak, peak_src = peaks()
        # one launch = a sweep (fused) or a colour pass.  Fused sweeps of one V-cycle on the finest level: the first
        # pre-smoothing sweep starts from the zero correction (no read of e: -8 B/cell), the first post-smoothing sweep
        # adds the prolonged coarse correction on the fly (+1 B/cell); average over the 2*S sweeps.
        if args.smoother == 1:
            S2 = 2 * args.smooth
            alg_bytes_launch = ((sweep_b - 8) + (sweep_b + 1) + (S2 - 2) * sweep_b) / S2 * cells_local
        else:
            alg_bytes_launch = sweep_b / 2.0 * cells_local
        kdur = k_ms / max(k_launches, 1) * 1e-3
        achieved = alg_bytes_launch / kdur / 1e9 if k_launches else None
        vcycle_bytes = level_b * cells_local * sum(1.0 / 8 ** d for d in range(depths))
        copy_ms = ms_copy / e2e_steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "details": {"mg_depths": depths, "bottom_bicgstab_iterations": bottom_iters, "slab_planes_per_gpu": nzl},
            "roofline": {"bound": "hbm", "kernel": "finest-level GSRB " + ("fused red+black sweep" if args.smoother == 1 else "colour pass"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": profiled_traffic(args),
                         "traffic_source": "profiles/" + TRAFFIC_SOURCE + " (ncu --set full of this command: the finest-level sweeps of one "
                                           "V-cycle; quoted only when its kernel_source_sha16 matches this build, else null)",
                         "kernel_source_sha16": kernel_fingerprint(), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_launch,
                         "launches_timed": k_launches, "avg_launch_ms": kdur * 1e3,
                         "kernel_share_of_step": k_ms / ms_prof,
                         "share_measured_on": "second pass of the same K steps launched eagerly with per-launch CUDA events "
                                              "(%.3f ms/step; the headline pass replays CUDA graphs)" % (ms_prof / args.steps),
                         "vcycle_algorithmic_GBps": vcycle_bytes / (ms_step * 1e-3) / 1e9,
                         "vcycle_frac_of_peak": vcycle_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         # SURVEY 8(d) asks for the fraction of the nominal 8 TB/s as well as of the measured copy bandwidth
                         "frac_of_nominal_8000": achieved / 8000.0 if achieved else None,
                         "vcycle_frac_of_nominal_8000": vcycle_bytes / (ms_step * 1e-3) / 1e9 / 8000.0},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": cells_total * 8, "d2h_bytes_per_step": cells_total * 8,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "copies_only_ms_per_step": copy_ms,
                    "host_copy_GBps_per_gpu_each_direction": cells_local * 8 / (copy_ms * 1e-3) / 1e9,
                    "frac_of_copy_bound": copy_ms / (ms_e2e / e2e_steps),
                    "what": "every step: pinned-host residual -> HBM, setToZero + V-cycle, correction -> pinned host, through the C ABI; "
                            "two buffer pairs, the copies on their own streams so that upload i+1, V-cycle i and download i-1 overlap. "
                            "copies_only = the same copies without the V-cycle on all ranks at once: the PCIe / host-memory floor"},
            "gpu_launches": launches, "clocks": clocks,
            "breakdown_rank0": dict(breakdown, note="per V-cycle, eager profiling pass, CUDA events per category on rank 0"),
        }
        if halo_info:
            line["details"]["halo"] = halo_info
        if strong:
            line["strong"] = strong
        if alt:
            line["max_grid_size_64"] = alt
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"], parity = cpu_baseline_and_parity(args, pb, ms_step)
        if parity:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
        if parity and not parity["ok"]:
            raise SystemExit("bench.py: PARITY CHECK FAILED: " + json.dumps(parity))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # libraries (NCCL's version banner, torchrun) may write to fd 1: keep the real stdout for the ONE JSON line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=512, help="cells per side per GPU")
    ap.add_argument("--cpu-n", type=int, default=0, help="side of the CPU problem (default: the workload's own n)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling (config C3) sub-record")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the multi-rank vs one-GPU bit comparison")
    ap.add_argument("--no-alt", action="store_true", help="skip the max_grid_size = 64 sub-record")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (mgic_ctx_set_option), repeatable")
    ap.add_argument("--smooth", type=int, default=2, help="numMGsmooth (pre = post = bottom)")
    ap.add_argument("--box", type=int, default=32, help="max_grid_size (sets the MG depth, Factory.cpp:168-172)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, config C5): n^3 cells per GPU; strong (config C3): one n^3 domain over all GPUs")
    ap.add_argument("--keep-b", action="store_true", help="stream bCoef even though it is identically 1")
    ap.add_argument("--smoother", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fused-cfg", type=int, default=None, help="tile shape of the fused sweep (tuning)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="halo planes by NVLink peer stores or ncclSend/ncclRecv")
    ap.add_argument("--overlap-halo", action="store_true", help="exchange on a second stream while the interior planes are swept")
    ap.add_argument("--fused-min-cells", type=int, default=None, help="levels below this use the per-colour kernel (tuning)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
