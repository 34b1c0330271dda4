#!/usr/bin/env python3
"""bench.py -- throughput of the EMME eigen hot path on B200 (and of the reference on the host CPU).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--npoints 8192] [--mode scan|rows]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

Workload (config.workload): BASELINE.json configs[3], the synthetic grid sweep -- C1 physics
(input-example.json, method=eigen, omega_d_coeff=1.0) with `npoints` grid nodes, default 8192, the
largest single-GPU point of the sweep.  A STEP is one pass of the hot path = one Newton/secant
iterate of EigenSolver::newtonTraceSecantIteration: dense step (solve A X = A', delta = -1/tr X),
re-assembly of the dim x dim matrix A(omega + delta), secant quotient.  The timed region is K
iterates that follow a fresh seed at the reference's initial guess; should a point converge by the
reference's stop rule inside the region, the next scan point is seeded inside the region too (its
two assemblies are counted, nothing is skipped).

    metric  matrix_elements_per_s = dim^2 * (matrices assembled in the timed region) / time
    value   inputs resident in HBM, CUDA events on the launching stream, max over ranks
    e2e     same loop through the public API with HOST buffers: every step uploads the eta/g/bi
            tables from pinned memory, and downloads (omega, delta) and the assembled matrix
    N > 1   --mode scan (default, weak scaling): every rank iterates its own scan point (k_rho),
            no collective on the data path;  --mode rows (strong): one problem, work items dealt to
            ranks, one NCCL all-reduce per assembly, reported in the extra key "row_sharded"

Extra keys: roofline (kernel 1, FP64 pipe), roofline_dense (kernel 2), cpu_baseline (oracle/_ref on
the box's host cores, bounded sample), c1 (converged-eigenvalue time of input-example.json), sweep,
pic (row N4: input-example.json AS SHIPPED, method PIC -- BASELINE configs[0] -- marker-stages/s of
the fused stage kernel against the HBM roofline, the whole run through solve_once_pic with host
buffers, and the reference's PIC_State/Integrator on the host cores for a few steps).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C1_PATH = ROOT / "tests" / "golden" / "inputs" / "c1.json"
REF_DRIVER = ROOT / "oracle" / "_ref" / "ref_driver"
PIC_PATH = ROOT / "tests" / "golden" / "inputs" / "pic.json"
PIC_DRIVER = ROOT / "oracle" / "_ref" / "pic_driver"
# algorithmic HBM bytes of the PIC stage kernel per marker and Integrator::step (three stages):
# per stage load eta 8 + w 16 + A 16 + B 16 + v_para 8 + v_perp 8 + p_weight 8 = 80 B and store
# eta 8 + w 16 + A 16 + B 16 = 56 B; the stage-1 velocity is stored once (16 B) and loaded once (16 B)
PIC_BYTES_PER_MARKER_STEP = 3 * (80 + 56) + 32
FLOP_FIXED = 194 + 20 * 8      # fixed complex arithmetic + 8 transcendentals
FLOP_TRIP = 14                 # per Miller recurrence trip


def c1_text(npoints, k_rho=None):
    txt = C1_PATH.read_text()
    txt, n = re.subn(r'"npoints": 1024', f'"npoints": {npoints}', txt)
    assert n == 1
    if k_rho is not None:
        txt, n = re.subn(r'"k_rho": 0.3182', f'"k_rho": {k_rho!r}', txt)
        assert n == 1
    return txt


def algorithmic_flops(st):
    return st["evals"] * FLOP_FIXED + FLOP_TRIP * (st["fwd_trips"] + st["bwd_trips"])


def fp64_flops(st):
    """The same count without the Miller forward trips: the kernel runs that search in FP32."""
    return st["evals"] * FLOP_FIXED + FLOP_TRIP * st["bwd_trips"]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        smax = max((float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()),
                   default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- reference arm
def run_ref(*args, timeout=1800):
    r = subprocess.run([str(REF_DRIVER), *map(str, args)], capture_output=True, text=True,
                       timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_driver {args[0]} failed: {r.stdout} {r.stderr}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def reference_step(npoints, inp_path, omega, target_s=12.0, cache={}):
    """One bounded CPU sample of the step on the reference (oracle/_ref, all host cores):
    (i) the per-pair work of matrixAssembler for a strided subset of rows through the reference's
    own DedicatedThreadPool, scaled by pair count; (ii) its LAPACK zsysv call at n = min(dim, 2048),
    scaled by (dim/n)^3.  Returns (estimated seconds per full step, description)."""
    total_pairs = npoints * (npoints - 1) // 2
    if "rate" not in cache:                       # calibrate on a tiny sample once
        c = run_ref("time_rows", inp_path, omega.real, omega.imag, 0, max(npoints // 2, 1), 2)
        cache["rate"] = c["pairs"] / c["seconds"]
    want_pairs = max(cache["rate"] * target_s, npoints)
    nrows = int(max(2, min(npoints, round(want_pairs / (npoints / 2)))))
    stride = max(npoints // nrows, 1)
    a = run_ref("time_rows", inp_path, omega.real, omega.imag, 0, stride, nrows)
    cache["rate"] = a["pairs"] / a["seconds"]
    t_asm = a["seconds"] * total_pairs / a["pairs"]
    nd = min(npoints, 2048)
    if ("dense", nd) not in cache:
        cache[("dense", nd)] = run_ref("time_dense", nd, 2)["zsysv_s"]
    t_dense = cache[("dense", nd)] * (npoints / nd) ** 3
    desc = (f"reference (oracle/_ref, unmodified sources) on {a['threads']} host threads: "
            f"{a['pairs']} of {total_pairs} pairs (rows 0::{stride} x{nrows}) through its thread pool in "
            f"{a['seconds']:.2f} s, scaled by pair count -> {t_asm:.1f} s/assembly; zsysv n={nd} "
            f"{cache[('dense', nd)]:.2f} s scaled by (dim/n)^3 -> {t_dense:.1f} s")
    return t_asm + t_dense, a["threads"], desc, a["seconds"] + cache[("dense", nd)]


def bench_reference(args, rank, world):
    if rank != 0:
        return
    npoints = args.npoints
    tmp = Path(os.environ.get("TMPDIR", "/tmp")) / f"emme_bench_c1_n{npoints}.json"
    tmp.write_text(c1_text(npoints))
    if not REF_DRIVER.exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver missing (run __graft_entry__.build())"}))
        return
    omega = complex(-0.8, 0.25)
    for _ in range(args.warmup):
        reference_step(npoints, tmp, omega, target_s=3.0)
    t_est, spent, threads, desc = 0.0, 0.0, 0, ""
    for _ in range(args.steps):
        t, threads, desc, s = reference_step(npoints, tmp, omega)
        t_est += t
        spent += s
    dim = npoints
    value = dim * dim * args.steps / t_est
    line = {
        "impl": "reference", "metric": "matrix_elements_per_s", "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_est / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "iters_per_s": args.steps / t_est,
        "config": {"workload": f"sweep: C1 physics, npoints={npoints}, dim={dim}, one Newton/secant iterate per step",
                   "npoints": npoints, "dim": dim, "parallelism": f"{threads} host threads"},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": threads, "kind": "reference",
                         "sample": desc + f"; {spent:.1f} s of CPU wall time measured over {args.steps} steps"},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if PIC_DRIVER.exists():      # row N4: the reference's PIC method on the same host cores, 4 of 180 steps
        r = subprocess.run([str(PIC_DRIVER), "time", str(PIC_PATH), "1", "4"], capture_output=True, text=True,
                           timeout=600)
        if r.returncode == 0:
            c = json.loads(r.stdout.strip().splitlines()[-1])
            line["pic"] = {"workload": "input-example.json as shipped (method PIC), 4 of 180 steps",
                           "marker_stages_per_s": c["marker_stages_per_s"], "cores": c["threads"],
                           "markers": c["markers"], "seconds": c["seconds"]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- B200 arm
def timed_iterates(solver, inp, omega0, steps, tol, reseed_points, e2e=None):
    """Run `steps` Newton iterates; returns dict(assemblies, reseeds, stats sums)."""
    out = dict(assemblies=0, reseeds=0, flops=0.0, asm_ms=0.0, dense_ms=0.0, dense_flops=0.0, sym_steps=0)
    point = 0
    for _ in range(steps):
        if e2e is not None:
            e2e["upload"]()
        solver.newtonTraceSecantIteration()
        st = solver.stats()
        out["assemblies"] += 1
        out["flops"] += algorithmic_flops(st)
        out["flops_fp64"] = out.get("flops_fp64", 0.0) + fp64_flops(st)
        out["asm_ms"] += st["assemble_ms"]
        out["dense_ms"] += st["dense_ms"]
        out["dense_flops"] += st["dense_flops"]       # flops of the path that ran (4 dim^3 symmetric, 26/3 dim^3 LU)
        out["sym_steps"] = st["sym_steps"]
        out["last_stats"] = st
        if e2e is not None:
            e2e["download"]()
        if abs(solver.d_eigen_value) < abs(tol * solver.eigen_value):
            point += 1                                   # next scan point (src/main.cpp:263-302)
            reseed_points(point)
            solver.seed(solver.eigen_value)              # continuation from the converged omega
            out["assemblies"] += 2
            out["reseeds"] += 1
    return out


def bench_pic(device, hbm_peak):
    """Row N4: the PIC method of input-example.json as shipped (1024 cells x 1024 markers per cell,
    180 steps of 0.25) and a 16x larger marker count, where the state no longer fits the L2."""
    from emme_b200 import Input, pic
    inp = Input(PIC_PATH)
    p, mpc, nt, dt = pic.pic_params(inp)
    out = {"workload": "input-example.json as shipped: method PIC, 1024 cells x 1024 markers per cell, "
                       f"{nt} steps of {dt} (three Runge-Kutta stages each)", "dtype": "f64"}
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get("pic_stage_kernel", {})
    except Exception:
        traffic = {}
    for key, m, steps in (("default", mpc, nt), ("markers_x4", 4 * mpc, 40), ("markers_x16", 16 * mpc, 20)):
        markers = pic.load_markers(p, m * p.npoints, seed=1)
        s = pic.PIC_State.from_markers(p, *markers, device=device)
        s.step(dt, 3)                               # warm-up, graph capture
        ms = []
        for _ in range(3):
            s.step(dt, steps)
            ms.append(s.timing()[0])
        t = min(ms) * 1e-3
        n = s.marker_num()
        gbs = PIC_BYTES_PER_MARKER_STEP * n * steps / t / 1e9
        tr = traffic.get(str(n))
        out[key] = {"markers": n, "cells": p.npoints, "steps": steps, "ms_per_step": 1e3 * t / steps,
                    "us_per_stage": 1e6 * t / steps / 3, "marker_stages_per_s": 3.0 * n * steps / t,
                    "gpu_launches_per_step": 6,
                    "roofline": {"kernel": "pic_stage_kernel (+ pic_field_kernel)", "bound": "hbm", "achieved": gbs,
                                 "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 "algorithmic_bytes_per_marker_step": PIC_BYTES_PER_MARKER_STEP,
                                 "traffic": (tr["dram_bytes_read"] + tr["dram_bytes_write"]) if tr else None,
                                 "traffic_note": tr["note"] if tr else None}}
        s.close()
    # the whole default run through the public API with host buffers: marker loading on the host,
    # upload, 180 steps, download of the field history, diagnostics and eigenvalue
    pic.solve_once_pic(inp, seed=1, device=device)   # warm
    t0 = time.perf_counter()
    res = pic.solve_once_pic(inp, seed=1, device=device)
    t1 = time.perf_counter()
    n = mpc * p.npoints
    out["e2e"] = {"seconds": t1 - t0, "marker_stages_per_s": 3.0 * n * nt / (t1 - t0),
                  "h2d_bytes": 48 * n, "d2h_bytes": 16 * p.npoints * nt,
                  "eigenvalue": res["eigenvalue"], "breakdown": res["timing"],
                  "note": "solve_once_pic: std::mt19937 marker loading and v_perp sort on one host core, upload, "
                          "all steps, field history download, util::calculate_omega; the eigen method gives "
                          "omega = (-0.8235, 0.2585) for the same physics (the PIC frequency carries no sign)"}
    if PIC_DRIVER.exists():
        r = subprocess.run([str(PIC_DRIVER), "time", str(PIC_PATH), "1", "4"], capture_output=True, text=True,
                           timeout=600)
        if r.returncode == 0:
            c = json.loads(r.stdout.strip().splitlines()[-1])
            out["cpu_baseline"] = {"value": c["marker_stages_per_s"], "unit": "marker-stages/s",
                                   "cores": c["threads"], "kind": "reference",
                                   "sample": f"reference PIC_State + Integrator (oracle/_ref/pic_driver, unmodified "
                                             f"include/solver_pic.h) on {c['threads']} host threads: {c['steps']} of "
                                             f"{nt} steps of the same {c['markers']} markers in {c['seconds']:.2f} s",
                                   "est_s_per_run": c["seconds"] / c["steps"] * nt}
    return out


def bench_b200(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    from emme_b200 import EigenSolver, Input, capi, parallel, solve_once_eigen

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    lib = capi.load()
    npoints = args.npoints
    k_rho0 = 0.3182 + 0.004 * rank if world > 1 else 0.3182     # one scan point per rank
    inp = Input(text=c1_text(npoints, k_rho0 if world > 1 else None))
    tol = inp.number("iteration_precision")
    omega0 = inp.initial_guess()
    p, n = inp.params()
    dim = n
    solver = EigenSolver.from_input(inp, device=local_rank)
    ext = torch.cuda.ExternalStream(solver.stream(), device=local_rank)

    def reseed_points(point):
        nxt = Input(text=c1_text(npoints, k_rho0 + 0.0005 * point))
        solver_tables[:] = list(nxt.tables())
        pp, _ = nxt.params()
        capi.check(lib.emme_set_params(solver._h, pp))
        capi.check(lib.emme_set_tables(solver._h, *[t.ctypes.data for t in solver_tables]))

    solver_tables = list(inp.tables())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- warm-up: a separate solve (seed + W iterates), then a fresh seed for the timed region
    solver.seed(omega0 * 1.01)
    timed_iterates(solver, inp, omega0, max(args.warmup, 3), tol, reseed_points)
    reseed_points(0)
    solver.seed(omega0)
    l0 = solver.stats()["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    run = timed_iterates(solver, inp, omega0, args.steps, tol, reseed_points)
    ev1.record(ext)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    launches = solver.stats()["launches"] - l0
    assemblies = sum_over_ranks(run["assemblies"])
    value = dim * dim * assemblies / (ms * 1e-3)
    omega_dev = solver.eigen_value

    # ---- e2e: same loop through host buffers (pinned), copies inside the timed region
    pin_tab = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
    for t, src in zip(pin_tab, solver_tables):
        t.copy_(torch.from_numpy(src))
    pin_A = torch.empty((dim, dim, 2), dtype=torch.float64).pin_memory()
    host_state = np.zeros(4)

    def upload():
        for t, src in zip(pin_tab, solver_tables):
            t.numpy()[:] = src
        capi.check(lib.emme_set_tables(solver._h, *[t.data_ptr() for t in pin_tab]))

    def download():
        # eigen_matrix -> pinned host memory on the handle's copy stream: overlaps the next iterate
        # (the previous step's copy is waited for first, so one buffer is enough)
        capi.check(lib.emme_copy_wait(solver._h))
        capi.check(lib.emme_copy_matrix_async(solver._h, 0, pin_A.data_ptr()))
        host_state[:] = (solver.eigen_value.real, solver.eigen_value.imag,
                         solver.d_eigen_value.real, solver.d_eigen_value.imag)

    reseed_points(0)
    solver.seed(omega0)
    barrier()
    t0 = time.perf_counter()
    run2 = timed_iterates(solver, inp, omega0, args.steps, tol, reseed_points,
                          e2e={"upload": upload, "download": download})
    capi.check(lib.emme_copy_wait(solver._h))      # the last matrix has landed in host memory
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    e2e_s = max_over_ranks(t1 - t0)
    e2e_value = dim * dim * sum_over_ranks(run2["assemblies"]) / e2e_s
    h2d = 3 * 8 * n
    d2h = 16 * dim * dim + 32

    # ---- roofline of kernel 1 (FP64 pipe) and kernel 2
    peak_tf, nominal_mhz = capi.C.c_double(), capi.C.c_double()
    capi.check(lib.emme_fp64_peak(local_rank, capi.C.byref(peak_tf), capi.C.byref(nominal_mhz)))
    asm_tf = run["flops"] / (run["asm_ms"] * 1e-3) / 1e12
    dense_flops = run["dense_flops"] / args.steps
    dense_tf = run["dense_flops"] / (run["dense_ms"] * 1e-3) / 1e12 if run["dense_ms"] else 0.0
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        tr = json.loads((ROOT / "profiles" / "traffic.json").read_text())["assemble_kernel"].get(str(npoints))
        if tr:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass

    extra = {}
    if rank == 0 and not args.quick:
        # C1 itself (input-example.json edited per SURVEY 8d): converged-eigenvalue time on one GPU
        c1 = Input(C1_PATH)
        s1 = EigenSolver.from_input(c1, device=local_rank)
        solve_once_eigen(c1, c1.initial_guess(), solver=s1)         # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w1, its, _ = solve_once_eigen(c1, c1.initial_guess(), solver=s1)
        t1 = time.perf_counter()
        st1 = s1.stats()
        extra["c1"] = {"workload": "input-example.json (method=eigen, omega_d_coeff=1.0), N=1024",
                       "converged_eigenvalue_s": t1 - t0, "newton_iterates": len(its),
                       "omega": [w1.real, w1.imag], "assemble_ms": st1["assemble_ms"],
                       "dense_ms": st1["dense_ms"],
                       "reference_omega": [-0.8234840422998696, 0.25848499305912503]}
        s1.close()

    if rank == 0 and world == 1 and not args.quick:
        # BASELINE configs[3]: the grid-size sweep, one seed + two iterates per size
        sweep = []
        for nn in (512, 1024, 2048, 4096):
            si = Input(text=c1_text(nn))
            sv = EigenSolver.from_input(si, device=local_rank)
            sv.seed(omega0)
            sv.newtonTraceSecantIteration()
            sv.newtonTraceSecantIteration()
            ss_ = sv.stats()
            fl = algorithmic_flops(ss_)
            sweep.append({"npoints": nn, "assemble_ms": ss_["assemble_ms"], "dense_ms": ss_["dense_ms"],
                          "elements_per_s": nn * nn / ((ss_["assemble_ms"] + ss_["dense_ms"]) * 1e-3),
                          "assemble_tflops": fl / (ss_["assemble_ms"] * 1e-3) / 1e12,
                          "assemble_frac_of_fp64_peak": fl / (ss_["assemble_ms"] * 1e-3) / 1e12 / peak_tf.value})
            sv.close()
        extra["sweep"] = sweep

    if rank == 0 and world == 1 and not args.quick:
        extra["pic"] = bench_pic(local_rank, hbm_peak)

    row_sharded = None
    if world > 1 and args.mode_rows:
        inp_r = Input(text=c1_text(npoints))
        pr, nr = inp_r.params()
        ss = parallel.ShardedEigenSolver(pr, nr, *inp_r.tables(), device=local_rank,
                                         exchange=args.exchange)
        ss.seed(omega0 * 1.01)
        for _ in range(2):
            ss.newtonTraceSecantIteration()
        ss.seed(omega0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ss.newtonTraceSecantIteration()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dist.barrier()
        rs = max_over_ranks(t1 - t0)
        row_sharded = {"scaling": "strong", "value": dim * dim * args.steps / rs, "unit": "elements/s",
                       "ms_per_step": 1e3 * rs / args.steps, "omega": [ss.eigen_value.real, ss.eigen_value.imag],
                       "exchange": ("peer stores from inside the assembly kernel (CUDA IPC over NVLink) + barrier"
                                    if args.exchange == "p2p" else
                                    "NCCL all-reduce(sum) of disjoint shares, 16*dim^2 bytes per assembly")}
        ss.close()

    pic_sharded = None
    if world > 1 and not args.quick:
        # row N4 across GPUs (weak scaling): 4 x 1024 x 1024 markers per GPU in contiguous blocks,
        # one NCCL all-reduce of the density (2 * npoints doubles) per Runge-Kutta stage
        from emme_b200 import pic
        pinp = Input(PIC_PATH)
        pp, mpc, _, pdt = pic.pic_params(pinp)
        per_gpu = 4 * mpc * pp.npoints
        markers = pic.load_markers(pp, per_gpu * world, seed=1)
        sp = parallel.ShardedPIC(pp, markers, device=local_rank)
        del markers
        sp.step(pdt, 3)
        sp.synchronize()
        psteps = 40
        barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pext = torch.cuda.ExternalStream(sp.state.stream(), device=local_rank)
        pe0.record(pext)
        sp.step(pdt, psteps)
        pe1.record(pext)
        barrier()
        pms = max_over_ranks(pe0.elapsed_time(pe1))
        f_last = sp.current_field()
        pic_sharded = {"scaling": "weak", "markers_per_gpu": per_gpu, "markers": per_gpu * world,
                       "cells": pp.npoints, "steps": psteps, "ms_per_step": pms / psteps,
                       "marker_stages_per_s": 3.0 * per_gpu * world * psteps / (pms * 1e-3), "unit": "marker-stages/s",
                       "exchange": "one NCCL all-reduce (sum) of 2*npoints doubles per stage on the handle's stream",
                       "field_rms": float(np.sqrt(np.mean(np.abs(f_last) ** 2)))}
        sp.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and REF_DRIVER.exists() and not args.quick:
        tmp = Path(os.environ.get("TMPDIR", "/tmp")) / f"emme_bench_c1_n{npoints}.json"
        tmp.write_text(c1_text(npoints))
        t_est, threads, desc, spent = reference_step(npoints, tmp, omega0, target_s=15.0)
        cpu_baseline = {"value": dim * dim / t_est, "unit": "elements/s", "cores": threads,
                        "kind": "reference", "sample": desc, "est_s_per_step": t_est}

    if rank == 0:
        st = run["last_stats"]
        line = {
            "metric": "matrix_elements_per_s", "value": value, "unit": "elements/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"sweep: C1 physics (input-example.json, method=eigen, omega_d_coeff=1.0), "
                                   f"npoints={npoints}, dim={dim}; step = one Newton/secant iterate "
                                   f"(dense step + assembly + secant)",
                       "npoints": npoints, "dim": dim,
                       "parallelism": "1 GPU" if world == 1 else f"scan-parallel: one k_rho point per GPU x{world}",
                       "l2": "each step rewrites >= 4 x 16*dim^2 bytes (4 GiB at npoints=8192), larger than L2",
                       "assemblies_timed": assemblies, "reseeds": run["reseeds"]},
            "iters_per_s": world * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "assemble_kernel<15> (kernel 1)", "bound": "fp64",
                         "achieved": asm_tf, "peak": peak_tf.value, "unit": "TFLOP/s",
                         "frac": asm_tf / peak_tf.value if peak_tf.value else None,
                         "peak_source": "DFMA micro-benchmark in this run (emme_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 FMA/clk x 2 x "
                                        f"{nominal_mhz.value:.0f} MHz = {148 * 64 * 2 * nominal_mhz.value / 1e6:.1f}",
                         "flops_per_launch": run["flops"] / args.steps,
                         "achieved_fp64_only": run["flops_fp64"] / (run["asm_ms"] * 1e-3) / 1e12,
                         "frac_fp64_only": (run["flops_fp64"] / (run["asm_ms"] * 1e-3) / 1e12 / peak_tf.value
                                            if peak_tf.value else None),
                         "note": "achieved = SURVEY 8d algorithmic flops (354/eval + 14 per Miller trip) / launch "
                                 "time; the forward (start-index) trips execute in FP32, achieved_fp64_only "
                                 "leaves them out",
                         "avg_launch_ms": run["asm_ms"] / args.steps,
                         "hbm_achieved_gbs": 16.0 * dim * dim / (run["asm_ms"] / args.steps * 1e-3) / 1e9,
                         "hbm_peak_gbs": hbm_peak, "traffic": traffic,
                         "algorithmic_bytes": 16 * dim * dim,
                         "evals": st["evals"], "fwd_trips": st["fwd_trips"], "bwd_trips": st["bwd_trips"]},
            "roofline_hbm": {"kernel": "assemble_kernel (kernel 1)", "bound": "hbm",
                             "achieved": 16.0 * dim * dim / (run["asm_ms"] / args.steps * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": 16.0 * dim * dim / (run["asm_ms"] / args.steps * 1e-3) / 1e9 / hbm_peak,
                             "traffic": traffic,
                             "note": "kernel 1 writes 16*dim^2 bytes once and reads 24*N bytes of tables: "
                                     "it is FP64-pipe bound (see roofline), HBM is idle"},
            "roofline_dense": {"kernel": "kernel 2: symmetric L D L^T + explicit-inverse trace (4 dim^3 flops) when "
                                         "sym_steps counts the step, LU + triangular solves (26/3 dim^3) otherwise",
                               "sym_steps": run["sym_steps"], "bound": "fp64",
                               "achieved": dense_tf, "peak": peak_tf.value, "unit": "TFLOP/s",
                               "frac": dense_tf / peak_tf.value if peak_tf.value else None,
                               "flops_per_step": dense_flops, "avg_ms": run["dense_ms"] / args.steps},
            "omega": [omega_dev.real, omega_dev.imag],
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if row_sharded:
            line["row_sharded"] = row_sharded
        if pic_sharded:
            line["pic_sharded"] = pic_sharded
        line.update(extra)
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--npoints", type=int, default=8192)
    ap.add_argument("--mode", default="scan", choices=["scan", "rows"])
    ap.add_argument("--quick", action="store_true", help="skip the cpu_baseline and C1 extras")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "allreduce"],
                    help="row-sharded mode: fused peer stores or NCCL all-reduce")
    args = ap.parse_args()
    args.mode_rows = True
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        bench_reference(args, rank, world)
    else:
        bench_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
