#!/usr/bin/env python3
"""Golden vectors of the PIC method (row N4) from the UNMODIFIED reference.

Run in the build container only: needs oracle/_ref/pic_driver (oracle/Makefile `make ref`, the
reference's include/solver_pic.h with a reproducible seed, see oracle/pic_driver.cpp).  Writes
tests/golden/pic_<case>.npz: the loaded markers, the derived tables, the field after every
Integrator::step, the final marker state and util::calculate_omega of the run; and
tests/golden/oscillator_rk3.bin: the Integrator template on x'' = -x (fixed and adaptive steps).
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
DRIVER = ROOT / "oracle" / "_ref" / "pic_driver"

# case -> (input file, seed, steps)
CASES = {
    "n32": ("pic_n32.json", 1234, 12),
    "n32_noswitch": ("pic_n32_noswitch.json", 99, 10),
    "n64_wb": ("pic_n64_wb.json", 7, 8),
    "n32_long": ("pic_n32.json", 5, 60),
}


def read_dump(path):
    b = Path(path).read_bytes()
    n, nf, nt = (int(v) for v in np.frombuffer(b, np.uint64, 3))
    o = 24

    def take(count, dtype=np.float64):
        nonlocal o
        a = np.frombuffer(b, dtype, count, o).copy()
        o += a.nbytes
        return a

    d = {"eta": take(n), "v_para": take(n), "v_perp": take(n), "weight": take(n, np.complex128),
         "omega_dv": take(n), "omega_st": take(n), "p_weight": take(n), "coef": take(nf),
         "fields": take(nf * nt, np.complex128).reshape(nt, nf), "eta_final": take(n),
         "weight_final": take(n, np.complex128), "omega": take(1, np.complex128)}
    assert o == len(b)
    return d


def main():
    if not DRIVER.exists():
        sys.exit("build oracle/_ref/pic_driver first (make -C oracle ref)")
    for case, (inp, seed, steps) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            out = Path(tmp) / "dump.bin"
            subprocess.run([str(DRIVER), "run", str(HERE / "inputs" / inp), str(seed), str(steps), str(out)],
                           check=True, cwd=tmp)
            d = read_dump(out)
        if case == "n32_long":   # only what the calculate_omega / growth checks need
            d = {k: d[k] for k in ("fields", "omega")}
        np.savez_compressed(HERE / f"pic_{case}.npz", seed=seed, steps=steps, input=inp, **d)
        print(case, {k: v.shape for k, v in d.items()}, d["omega"])
    # the reference's Integrator template on the oscillator of test/test_integrator.cpp
    subprocess.run([str(DRIVER), "oscillator", str(HERE / "oscillator_rk3.bin")], check=True)


if __name__ == "__main__":
    main()
