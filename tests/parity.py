"""Parity metric between an assembled matrix and the reference's (or the oracle's).

north_star: matrix entries within 1e-10 relative.  Measured fact (tests/golden/
rounding_floor.json, produced by make_goldens.py --floor): the REFERENCE ITSELF, rebuilt with
FMA contraction, deviates from its default build by up to 2.7e-10 relative on the tiniest
entries (|a| ~ 1e-11 .. 1e-7, i.e. 1e-5 .. 1e-9 of the quadrature's own absolute accuracy
`integration_accuracy`), because those entries are what is left after panel sums 1e5..1e9
times larger cancel.  The checker therefore reports three things and asserts all of them:

  * strict  : fraction of entries with |d| <= rtol*|a|            (reported; >= min_strict)
  * floor   : every entry satisfies |d| <= rtol*|a| + eps_floor*max|A_block|
              with eps_floor = 2^-52 -- one ulp of the block's largest entry, below which no
              consumer of A (the LU solve) can see a difference
  * flips   : no entry is off by more than 1e-8*|a| + 64 ulp(max|A_block|) (a changed
              accept/bisect decision of the adaptive quadrature moves an entry by ~1e-7
              absolute)
"""
import numpy as np

EPS = 2.0 ** -52


def blocks(A, em):
    if not em:
        return [A]
    n = A.shape[0] // 2
    return [A[:n, :n], A[:n, n:], A[n:, :n], A[n:, n:]]


def compare(A, ref, em=False, rtol=1e-10):
    out = dict(entries=0, strict_ok=0, floor_viol=0, max_rel=0.0, max_abs=0.0, max_abs_over_max=0.0,
               median_rel=0.0)
    rels = []
    for a, r in zip(blocks(A, em), blocks(ref, em)):
        d = np.abs(a - r)
        mag = np.abs(r)
        scale = mag.max() if mag.size else 0.0
        nz = mag > 0
        rel = np.zeros_like(d)
        rel[nz] = d[nz] / mag[nz]
        out["entries"] += int(d.size)
        out["strict_ok"] += int((d <= rtol * mag).sum())
        out["floor_viol"] += int((d > rtol * mag + EPS * scale).sum())
        out["flips"] = out.get("flips", 0) + int((d > 1e-8 * mag + 64 * EPS * scale).sum())
        out["max_rel"] = max(out["max_rel"], float(rel.max()))
        out["max_abs"] = max(out["max_abs"], float(d.max()))
        if scale > 0:
            out["max_abs_over_max"] = max(out["max_abs_over_max"], float(d.max() / scale))
        rels.append(rel[nz].ravel())
    allrel = np.concatenate(rels) if rels else np.zeros(0)
    out["median_rel"] = float(np.median(allrel)) if allrel.size else 0.0
    out["strict_frac"] = out["strict_ok"] / max(out["entries"], 1)
    out["n_rel_gt_1e9"] = int((allrel > 1e-9).sum())
    return out


def assert_parity(A, ref, em=False, rtol=1e-10, min_strict=0.90, label=""):
    assert A.shape == ref.shape
    assert np.isfinite(A.view(np.float64)).all(), f"{label}: non-finite entries"
    c = compare(A, ref, em=em, rtol=rtol)
    msg = f"{label}: {c}"
    assert c["floor_viol"] == 0, "entries beyond rtol*|a| + ulp(max|A|): " + msg
    assert c["strict_frac"] >= min_strict, "too few entries within strict rtol: " + msg
    assert c["flips"] == 0, "adaptive-tree flip suspected: " + msg
    return c
