#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Runs oracle/_ref/ref_driver (built by `make -C oracle ref` from /root/reference, build
container only) on tests/golden/inputs/*.json and stores

  golden.json            KAT values of the leaf numerics, derived constants, Newton iterate
                         lists (17 significant digits) for every case
  tables_<case>.npy      (3, N) float64: eta_i, g_integration_f(eta_i), bi(eta_i)
  A_<case>.npy           full A(omega0) (complex128) for the small cases
  rows_<case>.npz        selected rows of A(omega0) for the full-size cases C1 / C3

Usage: python tests/golden/make_goldens.py [--full]   (--full adds the N=1024 cases, ~10 min)
"""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
DRV = ROOT / "oracle" / "_ref" / "ref_driver"

SMALL = ["c1_n32", "c1_n64", "c1_n128", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64", "c1_cyl_n64",
         "c1_tmd_n64", "c1_cylold_n64", "c3_n32", "c3_n64", "c3_n128"]
NEWTON = ["c1_n64", "c1_n128", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64", "c3_n128"]
TABLES_ONLY = ["c1", "c3", "c1_n256", "c1_n512", "c3_n256", "c1_em_n128"]
FULL = ["c1", "c3"]
FULL_ROWS = [0, 1, 5, 100, 511, 1000, 1023]


def run(*args):
    r = subprocess.run([str(DRV), *map(str, args)], capture_output=True, text=True, check=True)
    return r.stdout


def last_json(out):
    """The Cylinder constructor prints shat_coeff to stdout (src/Parameters.cpp:397) before
    our JSON line; take the last line."""
    return json.loads(out.strip().splitlines()[-1])


def omega0(case):
    # the stellarator file spells numbers like 1.e-5, which Python's json rejects
    import re
    m = re.search(r'"initial_guess"\s*:\s*\[([^\]]*)\]', (HERE / "inputs" / f"{case}.json").read_text())
    return [float(x) for x in m.group(1).split(",")]


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
        elif t[0] == "TIME":
            for k, v in zip(t[1::2], t[2::2]):
                times[k] = float(v)
    return dict(seed=seed, iterates=iters, final=final, times=times)


def rounding_floor():
    """Self-consistency of the REFERENCE under a different rounding of the same arithmetic:
    ref_driver rebuilt with -march=native -ffp-contract=fast (FMA contraction) versus the
    default build.  The per-entry relative deviation of the tiniest entries (which are the
    result of cancellation between panel sums ~1e5..1e9 times larger) exceeds 1e-10 on some
    cases; tests use these numbers to tell rounding from real disagreement."""
    import os
    blas = next((Path("/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs")
                 ).glob("libopenblasp-*.so"))
    ref = Path("/root/reference")
    with tempfile.TemporaryDirectory() as td:
        exe = f"{td}/ref_driver_fma"
        srcs = [str(ref / "src" / n) for n in ("Parameters.cpp", "functions.cpp", "JsonParser.cpp",
                                                "singularity_handler.cpp", "Timer.cpp")]
        subprocess.run(["/usr/bin/g++", "-O3", "-std=c++20", "-march=native", "-ffp-contract=fast",
                        "-DEMME_EXPRESSION_TEMPLATE", "-DMULTI_THREAD", f"-I{ref}/include",
                        f"-I{ROOT}/oracle/shim", "-o", exe, str(ROOT / "oracle" / "ref_driver.cpp"), *srcs,
                        str(blas), f"-Wl,--disable-new-dtags,-rpath,{blas.parent}", "-lpthread"], check=True)
        out = {}
        for case in SMALL:
            w = omega0(case)
            subprocess.run([exe, "assemble", str(HERE / "inputs" / f"{case}.json"), repr(w[0]), repr(w[1]),
                            f"{td}/a.bin"], check=True, capture_output=True)
            refA = np.load(HERE / f"A_{case}.npy")
            A = np.fromfile(f"{td}/a.bin", dtype=np.complex128).reshape(refA.shape)
            d = np.abs(A - refA)
            m = np.abs(refA) > 0
            rel = d[m] / np.abs(refA[m])
            out[case] = dict(max_abs=float(d.max()), max_abs_over_maxA=float(d.max() / np.abs(refA).max()),
                             max_rel=float(rel.max()), median_rel=float(np.median(rel)),
                             n_rel_gt_1e10=int((rel > 1e-10).sum()), n_rel_gt_1e9=int((rel > 1e-9).sum()),
                             entries=int(m.sum()))
            print("floor", case, out[case])
    (HERE / "rounding_floor.json").write_text(json.dumps(out, indent=1))


NEWTON_QR = ["c1_n64", "c1_n128", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64"]


def newton_qr():
    """Iterate lists of the reference's OTHER iteration_method (newtonQRSecantIteration,
    include/solver.h:210-383): the same inputs with "iteration_method": "QRSecant"."""
    gpath = HERE / "golden.json"
    G = json.loads(gpath.read_text())
    G["newton_qr"] = {}
    with tempfile.TemporaryDirectory() as td:
        for case in NEWTON_QR:
            txt = (HERE / "inputs" / f"{case}.json").read_text()
            assert '"TraceSecant"' in txt
            Path(f"{td}/in.json").write_text(txt.replace('"TraceSecant"', '"QRSecant"'))
            G["newton_qr"][case] = parse_newton(run("newton", f"{td}/in.json"))
            print(case, "newton_qr", G["newton_qr"][case]["final"])
    gpath.write_text(json.dumps(G, indent=1))


def main():
    if "--floor" in sys.argv:
        return rounding_floor()
    if "--qr" in sys.argv:
        return newton_qr()
    full = "--full" in sys.argv
    gpath = HERE / "golden.json"
    G = json.loads(gpath.read_text()) if gpath.exists() else {}
    G["kat"] = json.loads(run("kat"))
    G.setdefault("tables", {})
    G.setdefault("assemble", {})
    G.setdefault("newton", {})
    with tempfile.TemporaryDirectory() as td:
        for case in SMALL + TABLES_ONLY:
            inp = HERE / "inputs" / f"{case}.json"
            out = run("tables", inp, f"{td}/t.bin")
            G["tables"][case] = last_json(out)
            np.save(HERE / f"tables_{case}.npy", np.fromfile(f"{td}/t.bin").reshape(3, -1))
        for case in SMALL:
            inp = HERE / "inputs" / f"{case}.json"
            w = omega0(case)
            out = run("assemble", inp, repr(w[0]), repr(w[1]), f"{td}/a.bin")
            info = last_json(out)
            A = np.fromfile(f"{td}/a.bin", dtype=np.complex128).reshape(info["dim"], info["dim"])
            np.save(HERE / f"A_{case}.npy", A)
            G["assemble"][case] = dict(omega=w, dim=info["dim"])
            print(case, "assembled", info)
        for case in NEWTON:
            G["newton"][case] = parse_newton(run("newton", HERE / "inputs" / f"{case}.json"))
            print(case, "newton", G["newton"][case]["final"])
        if full:
            for case in FULL:
                inp = HERE / "inputs" / f"{case}.json"
                w = omega0(case)
                info = last_json(run("assemble", inp, repr(w[0]), repr(w[1]), f"{td}/a.bin"))
                dim = info["dim"]
                A = np.fromfile(f"{td}/a.bin", dtype=np.complex128).reshape(dim, dim)
                rows = sorted(set(FULL_ROWS + ([r + dim // 2 for r in FULL_ROWS] if dim > 1024 else [])))
                absA = np.abs(A)
                np.savez_compressed(HERE / f"rows_{case}.npz", rows=np.array(rows), data=A[rows],
                                    fro=np.linalg.norm(A), maxabs=absA.max(),
                                    colsum=A.sum(axis=0), diag=np.diag(A))
                G["assemble"][case] = dict(omega=w, dim=dim, cpu_assemble_s=info["assemble_s"],
                                           threads=info["threads"])
                print(case, "assembled", info)
                G["newton"][case] = parse_newton(run("newton", inp))
                print(case, "newton", G["newton"][case])
    gpath.write_text(json.dumps(G, indent=1))


if __name__ == "__main__":
    main()
