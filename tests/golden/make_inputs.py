#!/usr/bin/env python3
"""Derive the parity inputs from the reference's two example files by TEXTUAL edits.

Run in the build container only (needs /root/reference); the produced JSON files are
committed under tests/golden/inputs/.  Edits follow SURVEY.md section 8d:

* C1  = input-example.json with "method": "eigen" and "omega_d_coeff": 1.0
  (the shipped file selects the PIC method and a scan object).
* C3  = input-stellarator-example.json plus the seven keys that src/main.cpp:187,41 and
  the Parameters constructor (src/Parameters.cpp:41,48,60,61,65-66) require.
* *_nXXX variants only change "npoints" (and nothing else).

The reference's lexer classifies a number as FLOAT only if it contains '.'
(src/JsonParser.cpp:436-443), so number spellings are preserved verbatim.
"""
import re
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "inputs"


def sub1(text, pattern, repl):
    new, n = re.subn(pattern, repl, text, count=1)
    assert n == 1, pattern
    return new


def main():
    OUT.mkdir(exist_ok=True)
    ex = (REF / "input-example.json").read_text()
    c1 = sub1(ex, r'"method": "PIC"', '"method": "eigen"')
    c1 = sub1(c1, r'"omega_d_coeff":\{[^}]*\}', '"omega_d_coeff": 1.0')
    (OUT / "c1.json").write_text(c1)
    # the shipped scan form, method switched to eigen only (11 continuation-chained points)
    (OUT / "c1_scan.json").write_text(sub1(ex, r'"method": "PIC"', '"method": "eigen"'))
    for n in (32, 64, 128, 256, 512):
        (OUT / f"c1_n{n}.json").write_text(sub1(c1, r'"npoints": 1024', f'"npoints": {n}'))
    # GK31 variant of C1 (same physics, other quadrature order / tolerances)
    g31 = sub1(c1, r'"integration_start_points": 15', '"integration_start_points": 31')
    (OUT / "c1_gk31_n128.json").write_text(sub1(g31, r'"npoints": 1024', '"npoints": 128'))
    # electromagnetic tokamak (beta_e != 0 -> dim = 2N, three integrals per pair)
    em = sub1(c1, r'"beta_e": 0.00', '"beta_e": 0.02')
    for n in (64, 128):
        (OUT / f"c1_em_n{n}.json").write_text(sub1(em, r'"npoints": 1024', f'"npoints": {n}'))
    # positive real frequency start (exercises omi = -sign(Re omega) = -1)
    pos = sub1(c1, r'"initial_guess": \[-0.8, 0.25\]', '"initial_guess": [0.8, 0.25]')
    (OUT / "c1_pos_n64.json").write_text(sub1(pos, r'"npoints": 1024', '"npoints": 64'))
    # other geometries of Parameters::generate (src/Parameters.cpp:18-31)
    for conf, tag in (("cylinder", "cyl"), ("taloyMagneticDrift", "tmd"), ("cylinder old", "cylold")):
        t = sub1(c1, r'"conf": "tokamak"', f'"conf": "{conf}"')
        (OUT / f"c1_{tag}_n64.json").write_text(sub1(t, r'"npoints": 1024', '"npoints": 64'))

    # PIC method (row N4): the shipped file with the scan object collapsed, small meshes
    pic = sub1(ex, r'"omega_d_coeff":\{[^}]*\}', '"omega_d_coeff": 1.0')
    (OUT / "pic.json").write_text(pic)
    p32 = sub1(sub1(pic, r'"npoints": 1024', '"npoints": 32'), r'"marker_per_cell":1024', '"marker_per_cell":32')
    (OUT / "pic_n32.json").write_text(p32)
    (OUT / "pic_n32_noswitch.json").write_text(
        sub1(p32, r'"drift_center_transformation_switch":true', '"drift_center_transformation_switch":false'))
    wb = sub1(sub1(pic, r'"npoints": 1024', '"npoints": 64'), r'"marker_per_cell":1024', '"marker_per_cell":24')
    wb = sub1(wb, r'"water_bag_weight_vpara": 1.0', '"water_bag_weight_vpara": 0.5')
    wb = sub1(wb, r'"water_bag_weight_vperp": 1.0', '"water_bag_weight_vperp": 1.5')
    (OUT / "pic_n64_wb.json").write_text(wb)

    st = (REF / "input-stellarator-example.json").read_text()
    extra = ('    "method":"eigen",\n    "iteration_method":"TraceSecant",\n    "epsilon_r":0.0,\n'
             '    "omega_d_coeff":1.0,\n    "water_bag_weight_vpara":1.0,\n'
             '    "water_bag_weight_vperp":1.0,\n    "drift_center_transformation_switch":true,\n')
    c3 = sub1(st, r'\{\n', '{\n' + extra)
    (OUT / "c3.json").write_text(c3)
    for n in (32, 64, 128, 256):
        (OUT / f"c3_n{n}.json").write_text(sub1(c3, r'"npoints":1024', f'"npoints":{n}'))
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    sys.exit(main())
