#!/usr/bin/env python3
"""bench.py -- throughput of the EMME eigen hot path on B200 (and of the reference on the host CPU).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--npoints 8192] [--quick]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

Workload (config.workload): BASELINE.json configs[3], the synthetic grid sweep -- C1 physics
(input-example.json, method=eigen, omega_d_coeff=1.0) with `npoints` grid nodes, default 8192, the
largest single-GPU point of the sweep.  A STEP is one pass of the hot path = one Newton/secant
iterate of EigenSolver::newtonTraceSecantIteration: dense step (solve A X = A', delta = -1/tr X),
re-assembly of the dim x dim matrix A(omega + delta), secant quotient.  The timed region is K
iterates that follow a fresh seed at the reference's initial guess; when a point converges by the
reference's stop rule inside the region, the next scan point is seeded inside the region as well
(its two assemblies cost time and are NOT counted as work: `value` counts iterates only, in both
arms, so the seeding overhead is charged against the B200 arm alone).

    metric  matrix_elements_per_s = dim^2 * (Newton iterates in the timed region) / time
    value   inputs resident in HBM, CUDA events on the launching stream, max over ranks
    e2e     same loop through the public API with HOST buffers: every step uploads the eta/g/bi
            tables from pinned memory and downloads (omega, delta); every converged point
            downloads its eigen_matrix (what solve_once_eigen writes to eigenMatrics/*.bin)
    N > 1   scan-parallel (weak scaling): every rank iterates its own scan point (k_rho), no
            collective on the data path

Exactly ONE JSON line is printed.  The headline is complete before any extra leg starts; every extra
is fenced (a failing extra records {"error": ...}), and if the run ends before the last leg has
finished (a hang -> watchdog, an uncaught exception, SIGTERM) the line is printed anyway with the
headline and the legs that did finish ("incomplete": ...).
Extras: c1 / c3 (converged-eigenvalue time of input-example.json and of the stellarator case, with
the reference's full solve timed in the same run), c5 (64-wavenumber scan over the ranks), sweep
(512 .. 8192, row-sharded over the ranks when N > 1), row_sharded (one N=8192 problem over N GPUs:
pair-sharded assembly + column-sharded dense step over NVLink peer memory), pic / pic_sharded
(row N4: input-example.json as shipped, method PIC).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from emme_b200 import workloads  # noqa: E402  (pure python: input texts of the configs)

REF_DRIVER = ROOT / "oracle" / "_ref" / "ref_driver"
PIC_PATH = workloads.PIC_PATH
PIC_DRIVER = ROOT / "oracle" / "_ref" / "pic_driver"
GOLDEN = ROOT / "tests" / "golden"
# algorithmic HBM bytes of the PIC stage kernel per marker and Integrator::step (three stages):
# per stage load eta 8 + w 16 + A 16 + B 16 + v_para 8 + v_perp 8 + p_weight 8 = 80 B and store
# eta 8 + w 16 + A 16 + B 16 = 56 B; the stage-1 velocity is stored once (16 B) and loaded once (16 B)
PIC_BYTES_PER_MARKER_STEP = 3 * (80 + 56) + 32
FLOP_FIXED = 194 + 20 * 8      # fixed complex arithmetic + 8 transcendentals
FLOP_TRIP = 14                 # per Miller recurrence trip
T_START = time.time()
DEADLINE_S = float(os.environ.get("EMME_BENCH_DEADLINE_S", "720"))   # watchdog: the line is never lost


def c1_text(npoints, k_rho=None):
    return workloads.c4_text(npoints, k_rho)


def algorithmic_flops(st):
    return st["evals"] * FLOP_FIXED + FLOP_TRIP * (st["fwd_trips"] + st["bwd_trips"])


def fp64_flops(st):
    """The same count without the Miller forward trips: the kernel runs that search in FP32."""
    return st["evals"] * FLOP_FIXED + FLOP_TRIP * st["bwd_trips"]


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bench_config(npoints):
    """config of the JSON line: byte-identical in the B200 and the reference arm."""
    return {"workload": workloads.workload_name(npoints), "npoints": npoints, "dim": npoints,
            "value_counts": "dim^2 per Newton iterate; seeding assemblies of follow-up scan points are timed, not counted",
            "l2": "each step rewrites >= 4 x 16*dim^2 bytes (4 GiB at npoints=8192), larger than L2"}


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
def ref_env():
    """The reference's pool sizes itself with hardware_concurrency(); its LAPACK (OpenBLAS) follows
    OMP_NUM_THREADS / OPENBLAS_NUM_THREADS, which torchrun exports as 1: give it every host core."""
    env = dict(os.environ)
    n = str(host_threads())
    env["OMP_NUM_THREADS"] = n
    env["OPENBLAS_NUM_THREADS"] = n
    env["MKL_NUM_THREADS"] = n
    return env


def run_ref(*args, timeout=1800, raw=False):
    r = subprocess.run([str(REF_DRIVER), *map(str, args)], capture_output=True, text=True,
                       timeout=timeout, env=ref_env())
    if r.returncode != 0:
        raise RuntimeError(f"ref_driver {args[0]} failed: {r.stdout[-300:]} {r.stderr[-300:]}")
    return r.stdout if raw else json.loads(r.stdout.strip().splitlines()[-1])


def parse_newton(out):
    seed, iters, final, times = None, [], None, {}
    for line in out.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "SEED":
            seed = [float(x) for x in t[1:5]]
        elif t[0] == "ITER":
            iters.append([float(x) for x in t[2:6]])
        elif t[0] == "FINAL":
            final = [float(t[1]), float(t[2]), int(t[3])]
        elif t[0] == "TIME" and t[1] == "initial":
            for k, v in zip(t[1::2], t[2::2]):
                times[k] = float(v)
    return dict(seed=seed, iterates=iters, final=final, times=times)


def reference_step(npoints, inp_path, omega, target_s=12.0, cache={}):
    """One bounded CPU sample of the step on the reference (oracle/_ref, all host cores):
    (i) the per-pair work of matrixAssembler for a strided subset of rows through the reference's
    own DedicatedThreadPool, scaled by pair count; (ii) its LAPACK zsysv call at n = min(dim, 2048),
    scaled by (dim/n)^3.  Returns (estimated seconds per full step, threads, description, seconds
    actually spent)."""
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
    desc = (f"reference (oracle/_ref, unmodified sources) on {a['threads']} host threads "
            f"(LAPACK threads {host_threads()}): "
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
    omega = complex(*workloads.C1_START)
    for _ in range(min(args.warmup, 1)):          # warm-up: page the binary and BLAS in, calibrate the rate
        reference_step(npoints, tmp, omega, target_s=3.0)
    # bounded sample per step (about 10 s of CPU work each), at most ~3 min for the whole run
    per_step_s = max(2.0, min(10.0, 170.0 / max(args.steps, 1)))
    t_est, spent, threads, desc = 0.0, 0.0, 0, ""
    for _ in range(args.steps):
        t, threads, desc, s = reference_step(npoints, tmp, omega, target_s=per_step_s)
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
        "config": bench_config(npoints),
        "parallelism": f"{threads} host threads (thread pool) / {host_threads()} LAPACK threads",
        "estimate": True,
        "measured_wall_s": spent,
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": threads, "kind": "reference",
                         "sample": desc + f"; {spent:.1f} s of CPU wall time measured over {args.steps} steps; "
                                          "ms_per_step is the EXTRAPOLATED time of a full step (a full N=8192 "
                                          "step needs ~4 min on 16 cores), measured_wall_s what was run"},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if PIC_DRIVER.exists() and not args.quick:      # row N4: the reference's PIC method on the same host cores
        try:
            r = subprocess.run([str(PIC_DRIVER), "time", str(PIC_PATH), "1", "4"], capture_output=True, text=True,
                               timeout=600, env=ref_env())
            if r.returncode == 0:
                c = json.loads(r.stdout.strip().splitlines()[-1])
                line["pic"] = {"workload": "input-example.json as shipped (method PIC), 4 of 180 steps",
                               "marker_stages_per_s": c["marker_stages_per_s"], "cores": c["threads"],
                               "markers": c["markers"], "seconds": c["seconds"]}
        except Exception as e:  # noqa: BLE001
            line["pic"] = {"error": repr(e)[:200]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- B200 arm
def timed_iterates(solver, steps, tol, reseed_points, e2e=None):
    """Run `steps` Newton iterates; when the reference's stop rule fires the next scan point is
    seeded (continuation from the converged omega, src/main.cpp:263-302)."""
    out = dict(assemblies=0, reseeds=0, flops=0.0, flops_fp64=0.0, asm_ms=0.0, dense_ms=0.0,
               dense_flops=0.0, sym_steps=0)
    point = 0
    for _ in range(steps):
        if e2e is not None:
            e2e["upload"]()
        solver.newtonTraceSecantIteration()
        st = solver.stats()
        out["assemblies"] += 1
        out["flops"] += algorithmic_flops(st)
        out["flops_fp64"] += fp64_flops(st)
        out["asm_ms"] += st["assemble_ms"]
        out["dense_ms"] += st["dense_ms"]
        out["dense_flops"] += st["dense_flops"]       # flops of the path that ran (4 dim^3 symmetric, 26/3 dim^3 LU)
        out["sym_steps"] = st["sym_steps"]
        out["last_stats"] = st
        if e2e is not None:
            e2e["download"]()
        if abs(solver.d_eigen_value) < abs(tol * solver.eigen_value):
            if e2e is not None:
                e2e["converged"]()
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
                  "note": "solve_once_pic: std::mt19937 marker loading on the host, upload, "
                          "all steps, field history download, util::calculate_omega; the eigen method gives "
                          "omega = (-0.8235, 0.2585) for the same physics (the PIC frequency carries no sign)"}
    if PIC_DRIVER.exists():
        r = subprocess.run([str(PIC_DRIVER), "time", str(PIC_PATH), "1", "4"], capture_output=True, text=True,
                           timeout=600, env=ref_env())
        if r.returncode == 0:
            c = json.loads(r.stdout.strip().splitlines()[-1])
            out["cpu_baseline"] = {"value": c["marker_stages_per_s"], "unit": "marker-stages/s",
                                   "cores": c["threads"], "kind": "reference",
                                   "sample": f"reference PIC_State + Integrator (oracle/_ref/pic_driver, unmodified "
                                             f"include/solver_pic.h) on {c['threads']} host threads: {c['steps']} of "
                                             f"{nt} steps of the same {c['markers']} markers in {c['seconds']:.2f} s",
                                   "est_s_per_run": c["seconds"] / c["steps"] * nt}
    return out


def bench_converged(name, path, golden_key, device, with_reference, peak_tf=None):
    """north_star target 1: converged-eigenvalue time of an input file through solve_once_eigen
    with HOST buffers (parse + tables + create + seed + iterates + eigen_matrix download), next to
    the reference's own full solve (ref_driver newton = its ctor + iteration loop) timed in the
    same run on all host cores, with omega parity asserted."""
    import numpy as np
    import torch
    from emme_b200 import EigenSolver, Input, solve_once_eigen
    inp = Input(path)
    s = EigenSolver.from_input(inp, device=device)
    solve_once_eigen(inp, inp.initial_guess(), solver=s)         # warm (graph capture, pool)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    w, its, _ = solve_once_eigen(inp, inp.initial_guess(), solver=s)
    t1 = time.perf_counter()
    st = s.stats()
    s.close()
    # cold: everything a caller of the reference's solve_once_eigen pays, host buffers both ways
    t2 = time.perf_counter()
    inp2 = Input(path)
    s2 = EigenSolver.from_input(inp2, device=device)
    w2, its2, _ = solve_once_eigen(inp2, inp2.initial_guess(), solver=s2)
    A = np.empty((s2.dim, s2.dim), dtype=np.complex128)
    s2._lib.emme_copy_matrix(s2._h, 0, A.ctypes.data)
    s2.close()
    t3 = time.perf_counter()
    out = {"workload": name, "npoints": inp.params()[1], "dim": s2.dim,
           "converged_eigenvalue_s": t1 - t0, "newton_iterates": len(its),
           "e2e_cold_s": t3 - t2,
           "e2e_cold_includes": "input.json parse, tables, emme_create (device buffers), seed, iterates, "
                                "eigen_matrix download to host memory, emme_destroy",
           "omega": [w.real, w.imag], "last_assemble_ms": st["assemble_ms"], "last_dense_ms": st["dense_ms"]}
    fl = algorithmic_flops(st)
    out["roofline"] = {"kernel": "assemble_kernel (kernel 1), last assembly of the solve", "bound": "fp64",
                       "achieved": fl / (st["assemble_ms"] * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                       "frac": (fl / (st["assemble_ms"] * 1e-3) / 1e12 / peak_tf) if peak_tf else None,
                       "evals": st["evals"], "integrals": st["integrals"]}
    d3 = float(s2.dim) ** 3
    out["roofline_dense"] = {"kernel": "kernel 2 (symmetric path, 4 dim^3 flops)", "bound": "fp64",
                             "achieved": st["dense_flops"] / (st["dense_ms"] * 1e-3) / 1e12 if st["dense_ms"] else None,
                             "peak": peak_tf, "unit": "TFLOP/s", "flops": st["dense_flops"], "dim3": d3}
    try:
        gold = json.loads((GOLDEN / "golden.json").read_text())["newton"][golden_key]["final"]
        out["golden_omega"] = gold[:2]
        out["golden_iterates"] = gold[2]
        out["rel_err_vs_golden"] = abs(w - complex(gold[0], gold[1])) / abs(complex(gold[0], gold[1]))
    except Exception:
        pass
    if with_reference and REF_DRIVER.exists():
        r = parse_newton(run_ref("newton", path, raw=True, timeout=900))
        wr = complex(r["final"][0], r["final"][1])
        t_ref = r["times"]["initial"] + r["times"]["iteration"]
        out["reference"] = {"omega": r["final"][:2], "iterates": r["final"][2], "seconds": t_ref,
                            "threads": int(r["times"].get("threads", 0)), "lapack_threads": host_threads(),
                            "kind": "reference (oracle/_ref/ref_driver newton: unmodified EigenSolver ctor + "
                                    "newtonTraceSecantIteration loop, its own stop rule)"}
        out["rel_err_vs_reference"] = abs(w - wr) / abs(wr)
        out["parity_ok"] = bool(out["rel_err_vs_reference"] <= 1e-8 and len(its) == r["final"][2])
        out["speedup_vs_reference"] = t_ref / (t1 - t0)
        out["speedup_vs_reference_e2e_cold"] = t_ref / (t3 - t2)
    return out


def bench_b200(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    from emme_b200 import EigenSolver, Input, capi, parallel

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
    state = {"line": None, "extras": {}, "printed_final": False, "emitted": False}
    import atexit
    import signal
    atexit.register(lambda: emit(False))            # an exception or sys.exit after the headline is known

    def on_term(signum, _frame):
        emit(False)
        os._exit(0 if state["emitted"] or rank != 0 else 1)
    signal.signal(signal.SIGTERM, on_term)

    def emit(final):
        """ONE JSON line on stdout, always: the complete line at the end of a normal run; if the run
        ends any other way (an extra leg that hangs -> watchdog, an uncaught exception, SIGTERM) the
        headline that was already known, with whatever extras had finished."""
        if rank != 0 or state["line"] is None or state["emitted"]:
            return
        state["emitted"] = True
        line = dict(state["line"])
        line.update(state["extras"])
        line["bench_wall_s"] = time.time() - T_START
        if not final:
            line["incomplete"] = "the run ended before every extra leg had finished; headline and finished legs only"
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()

    def watchdog():
        # never lose the line: if an extra hangs (a peer that died inside a collective), print what
        # is known and leave
        if not state["printed_final"]:
            state["extras"]["watchdog"] = f"an extra leg did not finish within {DEADLINE_S:.0f} s; line printed by the watchdog"
            state["printed_final"] = True
            emit(False)
            os._exit(0)

    wd = threading.Timer(max(30.0, DEADLINE_S - (time.time() - T_START)), watchdog)
    wd.daemon = True
    wd.start()

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
    timed_iterates(solver, max(args.warmup, 3), tol, reseed_points)
    reseed_points(0)
    solver.seed(omega0)
    l0 = solver.stats()["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    run = timed_iterates(solver, args.steps, tol, reseed_points)
    ev1.record(ext)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    launches = solver.stats()["launches"] - l0
    assemblies = sum_over_ranks(run["assemblies"])
    value = dim * dim * world * args.steps / (ms * 1e-3)
    omega_dev = solver.eigen_value

    # ---- e2e: same loop through host buffers (pinned), copies inside the timed region
    pin_tab = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
    for t, src in zip(pin_tab, solver_tables):
        t.copy_(torch.from_numpy(src))
    pin_A = torch.empty((dim, dim, 2), dtype=torch.float64).pin_memory()
    host_state = np.zeros(4)
    d2h = {"bytes": 0}

    def upload():
        for t, src in zip(pin_tab, solver_tables):
            t.numpy()[:] = src
        capi.check(lib.emme_set_tables(solver._h, *[t.data_ptr() for t in pin_tab]))

    def download():
        host_state[:] = (solver.eigen_value.real, solver.eigen_value.imag,
                         solver.d_eigen_value.real, solver.d_eigen_value.imag)
        d2h["bytes"] += 32

    def converged():
        # eigen_matrix of the converged point -> pinned host memory (what solve_once_eigen writes to
        # eigenMatrics/*.bin, src/main.cpp:59-63) on the handle's copy stream: overlaps the seeding of
        # the next point; the handle makes the assembly that overwrites the buffer wait for the copy
        capi.check(lib.emme_copy_wait(solver._h))
        capi.check(lib.emme_copy_matrix_async(solver._h, 0, pin_A.data_ptr()))
        d2h["bytes"] += 16 * dim * dim

    reseed_points(0)
    solver.seed(omega0)
    barrier()
    t0 = time.perf_counter()
    run2 = timed_iterates(solver, args.steps, tol, reseed_points,
                          e2e={"upload": upload, "download": download, "converged": converged})
    converged()                                    # the final state is read back like a finished solve
    capi.check(lib.emme_copy_wait(solver._h))      # the last matrix has landed in host memory
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    e2e_s = max_over_ranks(t1 - t0)
    e2e_value = dim * dim * world * args.steps / e2e_s
    h2d = 3 * 8 * n
    d2h_per_step = d2h["bytes"] / args.steps

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
    traffic, ncu_pipe = None, None
    try:
        tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())["assemble_kernel"].get(str(npoints))
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            ncu_pipe = tj.get("fp64_pipe_pct")
    except Exception:
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and REF_DRIVER.exists() and not args.quick:
        try:
            tmp = Path(os.environ.get("TMPDIR", "/tmp")) / f"emme_bench_c1_n{npoints}.json"
            tmp.write_text(c1_text(npoints))
            t_est, threads, desc, spent = reference_step(npoints, tmp, omega0, target_s=15.0)
            cpu_baseline = {"value": dim * dim / t_est, "unit": "elements/s", "cores": threads,
                            "kind": "reference", "sample": desc, "est_s_per_step": t_est}
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"error": repr(e)[:300]}

    if rank == 0:
        st = run["last_stats"]
        line = {
            "metric": "matrix_elements_per_s", "value": value, "unit": "elements/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(npoints),
            "parallelism": "1 GPU" if world == 1 else f"scan-parallel: one k_rho point per GPU x{world}, no data-path collective",
            "assemblies_timed": assemblies, "reseeds": run["reseeds"],
            "elements_assembled_per_s": dim * dim * assemblies / (ms * 1e-3),
            "iters_per_s": world * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h_per_step, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "note": "per step: tables uploaded from pinned memory, (omega, delta) read back; per converged "
                            "point and at the end: eigen_matrix (16*dim^2 bytes) to pinned host memory"},
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
                         "fp64_pipe_pct_ncu": ncu_pipe,
                         "note": "achieved = SURVEY 8d algorithmic flops (354/eval + 14 per Miller trip) / launch "
                                 "time; the forward (start-index) trips execute in FP32, achieved_fp64_only "
                                 "leaves them out; fp64_pipe_pct_ncu is the hardware counter "
                                 "(sm__inst_executed_pipe_fp64, profiles/traffic.json)",
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
        state["line"] = line
    solver.close()
    del pin_A
    if rank == 0:       # the headline is known: a copy for whoever watches the run (stderr, not the result line)
        sys.stderr.write("[bench] headline: " + json.dumps({k: state["line"][k] for k in
                         ("metric", "value", "unit", "n_gpus", "ms_per_step", "e2e")}) + "\n")
        sys.stderr.flush()

    # ------------------------------------------------------------------ extras, each fenced
    extras = state["extras"]

    def fenced(name, fn):
        t_leg = time.time()
        try:
            r = fn()
        except Exception as e:  # noqa: BLE001 - recorded, the line survives
            r = {"error": f"{type(e).__name__}: {e}"[:400]}
        if r is not None:
            if isinstance(r, dict):
                r["leg_wall_s"] = time.time() - t_leg
            extras[name] = r

    def time_left():
        return DEADLINE_S - (time.time() - T_START)

    def leg_c5():
        """BASELINE configs[4]: 64 independent wavenumbers (explicit starts) dealt round-robin to
        the ranks, N=1024; three points are checked against eigenvalues of the unmodified reference
        (tests/golden/c5.json, produced by tests/golden/make_c5_goldens.py)."""
        pts = workloads.c5_points(1024)
        texts = [t for _, _, t in pts]
        starts = [w for _, w, _ in pts]
        parallel.solve_scan_texts(texts[: 2 * world], starts[: 2 * world], device=local_rank)   # warm
        barrier()
        t0 = time.perf_counter()
        recs = parallel.solve_scan_texts(texts, starts, device=local_rank)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        secs = max_over_ranks(t1 - t0)
        if rank != 0:
            return None
        iters = sum(r.get("iterations", 0) for r in recs)
        out = {"workload": "C5: C1 physics, k_rho = 0.05 + 0.01 k, k = 0..63, omega0 = (-0.8, 0.25) k_rho/0.3182, "
                           "N=1024, independent points dealt round-robin to the ranks (solve_scan_texts)",
               "points": len(recs), "seconds": secs, "points_per_s": len(recs) / secs,
               "newton_iterates": iters, "iters_per_s": iters / secs,
               "matrix_elements_per_s": 1024 * 1024 * (iters + 2 * len(recs)) / secs,
               "converged": sum(1 for r in recs if r.get("converged")),
               "failed": sum(1 for r in recs if r.get("eigenvalue") == "NaN"),
               "scaling": "strong: the same 64 points over N GPUs", "n_gpus": world,
               "eigenvalues": [r.get("eigenvalue") for r in recs]}
        try:
            gold = json.loads((GOLDEN / "c5.json").read_text())["points"]
            chk = {}
            for k, g in gold.items():
                rec = recs[int(k)]
                if "final" not in g:          # the reference itself fails on this point: so must we
                    chk[k] = {"reference": "fails: " + g.get("error", "")[:80],
                              "ok": rec.get("eigenvalue") == "NaN", "ours": str(rec.get("reason", rec.get("eigenvalue")))[:120]}
                    continue
                wg = complex(g["final"][0], g["final"][1])
                w = rec["eigenvalue"]
                if w == "NaN":
                    chk[k] = {"ok": False, "ours": rec.get("reason")}
                    continue
                chk[k] = {"rel_err": abs(complex(*w) - wg) / abs(wg), "iterates": rec["iterations"],
                          "reference_iterates": g["final"][2]}
                chk[k]["ok"] = bool(chk[k]["rel_err"] <= 1e-8 and chk[k]["iterates"] == g["final"][2])
            out["vs_reference_goldens"] = chk
            out["parity_ok"] = all(c["ok"] for c in chk.values())
        except Exception as e:  # noqa: BLE001
            out["vs_reference_goldens"] = {"error": repr(e)[:200]}
        return out

    def leg_sweep():
        """BASELINE configs[3]: the grid-size sweep, one seed + two iterates per size; at N > 1 ONE
        problem per size is sharded over the ranks (pair-sharded assembly; column-sharded dense
        step from dim 4096 up)."""
        sweep = []
        for nn in workloads.C4_SIZES:
            if nn == npoints and world == 1:
                continue                       # the headline IS this point
            si = Input(text=c1_text(nn))
            if world == 1:
                sv = EigenSolver.from_input(si, device=local_rank)
            else:
                pr, nr = si.params()
                sv = parallel.ShardedEigenSolver(pr, nr, *si.tables(), device=local_rank)
            sv.seed(omega0)
            sv.newtonTraceSecantIteration()
            barrier()
            t0 = time.perf_counter()
            sv.newtonTraceSecantIteration()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            step_s = max_over_ranks(t1 - t0)
            ss_ = sv.stats()
            fl = sum_over_ranks(algorithmic_flops(ss_))
            asm_ms = max_over_ranks(ss_["assemble_ms"])
            rec = {"npoints": nn, "assemble_ms": asm_ms, "dense_ms": ss_["dense_ms"],
                   "step_ms": 1e3 * step_s, "elements_per_s": nn * nn / step_s,
                   "assemble_tflops": fl / (asm_ms * 1e-3) / 1e12,
                   "assemble_frac_of_fp64_peak": fl / (asm_ms * 1e-3) / 1e12 / (peak_tf.value * world),
                   "omega": [sv.eigen_value.real, sv.eigen_value.imag]}
            if world > 1:
                rec["dense_sharded"] = bool(getattr(sv, "dense_sharded", False))
            sweep.append(rec)
            sv.close()
        return sweep if rank == 0 else None

    def leg_row_sharded():
        inp_r = Input(text=c1_text(npoints))
        pr, nr = inp_r.params()
        ss = parallel.ShardedEigenSolver(pr, nr, *inp_r.tables(), device=local_rank)
        tolr = inp_r.number("iteration_precision")
        # parity first: seed + two iterates must reproduce a single-GPU solver bit for bit (the shards
        # are disjoint and every element sees the same operations in the same order for any number of
        # ranks); rank 0 runs the single-GPU solver, every rank compares with its result
        ss.seed(omega0 * 1.01)
        sharded_its = []
        for _ in range(2):
            ss.newtonTraceSecantIteration()
            sharded_its.append((ss.eigen_value, ss.d_eigen_value))
        single_its = [None]
        if rank == 0:
            s1 = EigenSolver.from_input(inp_r, device=local_rank)
            s1.seed(omega0 * 1.01)
            its = []
            for _ in range(2):
                s1.newtonTraceSecantIteration()
                its.append((s1.eigen_value, s1.d_eigen_value))
            s1.close()
            single_its = [its]
        dist.broadcast_object_list(single_its, src=0)
        bitwise = sum_over_ranks(1.0 if sharded_its == single_its[0] else 0.0) == world
        ss.seed(omega0)
        barrier()
        rsteps = min(args.steps, 20)
        t0 = time.perf_counter()
        asm = dns = it_s = 0.0
        reseeds = 0
        fb0 = ss.stats()["pivot_fallbacks"]
        for _ in range(rsteps):
            ti = time.perf_counter()
            ss.newtonTraceSecantIteration()
            it_s += time.perf_counter() - ti
            st_ = ss.stats()
            asm += st_["assemble_ms"]
            dns += st_["dense_ms"]
            if abs(ss.d_eigen_value) < abs(tolr * ss.eigen_value):
                ss.seed(ss.eigen_value)            # reseed on convergence, like timed_iterates
                reseeds += 1
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        dist.barrier()
        rs = max_over_ranks(t1 - t0)
        out = {"scaling": "strong", "value": dim * dim * rsteps / rs, "unit": "elements/s",
               "ms_per_step": 1e3 * rs / rsteps, "steps": rsteps, "reseeds": reseeds,
               "ms_per_iterate": 1e3 * max_over_ranks(it_s) / rsteps,
               "note": "ms_per_step includes the two seeding assemblies of every reseed; ms_per_iterate is the "
                       "host wall time of newtonTraceSecantIteration alone (dense step + assembly + secant)",
               "assemble_ms": max_over_ranks(asm / rsteps), "dense_ms": max_over_ranks(dns / rsteps),
               "pivot_fallbacks": ss.stats()["pivot_fallbacks"] - fb0,
               "bitwise_equal_to_single_gpu": bool(bitwise),
               "dense_sharded": bool(ss.dense_sharded),
               "omega": [ss.eigen_value.real, ss.eigen_value.imag],
               "exchange": "assembly: peer stores from inside the kernel (CUDA IPC over NVLink), device-side "
                           "epoch barriers; dense step: column-block-cyclic L D L^T with panel broadcast by "
                           "peer stores and flag signalling (no NCCL on the data path)"}
        ss.close()
        return out if rank == 0 else None

    def leg_pic_sharded():
        # row N4 across GPUs (weak scaling): 4 x 1024 x 1024 markers per GPU in contiguous blocks
        from emme_b200 import pic
        pinp = Input(PIC_PATH)
        pp, mpc, _, pdt = pic.pic_params(pinp)
        per_gpu = 4 * mpc * pp.npoints
        sp = parallel.ShardedPIC.from_seed(pp, per_gpu * world, seed=1, device=local_rank)
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
        out = {"scaling": "weak", "markers_per_gpu": per_gpu, "markers": per_gpu * world,
               "cells": pp.npoints, "steps": psteps, "ms_per_step": pms / psteps,
               "marker_stages_per_s": 3.0 * per_gpu * world * psteps / (pms * 1e-3), "unit": "marker-stages/s",
               "exchange": sp.exchange_description(),
               "field_rms": float(np.sqrt(np.mean(np.abs(f_last) ** 2)))}
        sp.close()
        return out if rank == 0 else None

    if not args.quick:
        if rank == 0 and world == 1:
            fenced("c1", lambda: bench_converged("input-example.json (method=eigen, omega_d_coeff=1.0), N=1024",
                                                 workloads.C1_PATH, "c1", local_rank, not args.no_reference,
                                                 peak_tf.value))
            fenced("c3", lambda: bench_converged("input-stellarator-example.json + the 7 missing keys (SURVEY 8d): "
                                                 "EM, GK31, N=1024, dim=2048", workloads.C3_PATH, "c3", local_rank,
                                                 not args.no_reference and time_left() > 300, peak_tf.value))
        fenced("c5", leg_c5)
        fenced("sweep", leg_sweep)
        if world > 1:
            fenced("row_sharded", leg_row_sharded)
            if time_left() > 150:
                fenced("pic_sharded", leg_pic_sharded)
        elif rank == 0:
            fenced("pic", lambda: bench_pic(local_rank, hbm_peak))

    state["printed_final"] = True
    wd.cancel()
    emit(True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--npoints", type=int, default=8192)
    ap.add_argument("--quick", action="store_true", help="headline only: skip cpu_baseline and the extras")
    ap.add_argument("--no-reference", action="store_true", help="skip the reference's full C1/C3 solves")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        bench_reference(args, rank, world)
    else:
        bench_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
