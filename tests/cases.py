"""Shared helpers: load a golden case (input file, reference tables, oracle parameter dict)."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def input_path(case):
    return GOLD / "inputs" / f"{case}.json"


def ref_tables(case):
    t = np.load(GOLD / f"tables_{case}.npy")
    return t[0].copy(), t[1].copy(), t[2].copy()


def ref_matrix(case):
    return np.load(GOLD / f"A_{case}.npy")


def oracle_params(p):
    """emme_params (ctypes, product ABI) -> dict for the oracle's emme_oracle_params."""
    d = p.as_dict()
    return dict(q=d["q"], R=d["R"], vt=d["vt"], tau=d["tau"], beta_e=d["beta_e"], eta_i=d["eta_i"],
                eta_e=d["eta_e"], omega_s_i=d["omega_s_i"], omega_s_e=d["omega_s_e"],
                omega_d_bar=d["omega_d_bar"], arc_coeff=d["arc_coeff"], tol=d["integration_precision"],
                prec=d["integration_accuracy"], maxdepth=d["integration_iteration_limit"],
                order=d["integration_start_points"])
