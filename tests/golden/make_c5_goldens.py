#!/usr/bin/env python3
"""Golden eigenvalues for config C5 (BASELINE configs[4]: 64 independent wavenumbers) from the
UNMODIFIED reference: oracle/_ref/ref_driver newton on the byte-identical input texts of
emme_b200/workloads.py::c5_point, once per file (the reference's own scan is a sequential
continuation chain, src/main.cpp:78,263,302).  Stores tests/golden/c5.json:
    {"npoints": N, "points": {"<k>": {"k_rho":, "omega0":, "iterates": [[wr, wi, dr, di] ...],
                                     "final": [wr, wi, n_iter]}}}
Usage: python tests/golden/make_c5_goldens.py [--npoints 1024] k [k ...]    (~95 s per point at
N=1024 on the 8-core build container)."""
import argparse
import json
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
from make_goldens import parse_newton, run  # noqa: E402

from emme_b200 import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npoints", type=int, default=1024)
    ap.add_argument("k", type=int, nargs="+")
    a = ap.parse_args()
    out_path = HERE / ("c5.json" if a.npoints == 1024 else f"c5_n{a.npoints}.json")
    G = json.loads(out_path.read_text()) if out_path.exists() else {"npoints": a.npoints, "points": {}}
    with tempfile.TemporaryDirectory() as td:
        for k in a.k:
            k_rho, w0, txt = workloads.c5_point(k, a.npoints)
            Path(f"{td}/in.json").write_text(txt)
            try:
                r = parse_newton(run("newton", f"{td}/in.json"))
                rec = {"k_rho": k_rho, "omega0": [w0.real, w0.imag], "iterates": r["iterates"],
                       "final": r["final"], "times": r["times"]}
            except Exception as e:  # noqa: BLE001 - the reference's failure IS the golden
                lines = (getattr(e, "stdout", "") or str(e)).splitlines()
                rec = {"k_rho": k_rho, "omega0": [w0.real, w0.imag],
                       "iterates": [[float(x) for x in ln.split()[2:6]] for ln in lines if ln.startswith("ITER")],
                       "error": next((ln[6:] for ln in lines if ln.startswith("ERROR")), "\n".join(lines)[-300:]),
                       "note": "the reference's scan records this point as {\"eigenvalue\": \"NaN\"} (src/main.cpp:311-318)"}
            G["points"][str(k)] = rec
            print(k, rec.get("final"), rec.get("error"), flush=True)
            out_path.write_text(json.dumps(G, indent=1))


if __name__ == "__main__":
    main()
