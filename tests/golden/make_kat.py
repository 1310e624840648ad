"""Regenerates tests/golden/kat_interference.json from the CPU oracle.

The values were first derived independently from the formulas cited in SURVEY.md section 8c
(interference.rs:197-353,380-400) and agree with the oracle to all printed digits; the file freezes
them so that a later edit of the oracle cannot silently move the target.
Run: python tests/golden/make_kat.py
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_lib as O  # noqa: E402

ic = O.OracleInterference(1.33)
lin_pts = [(500.0, 1.0), (300.0, 1.0), (300.0, 0.5), (100.0, 1.0), (400.0, 1.0), (250.0, math.sqrt(0.5)),
           (1000.0, 0.0), (737.0, 0.25), (1999.0, 0.9), (30.0, 0.05)]
gam_pts = [(500.0, 1.0), (300.0, 0.5), (737.0, 0.25), (1999.0, 0.9)]
kat = {
    "n_film": 1.33,
    "fresnel_normal": ic.fresnel(1.0)[2],
    "linear": [{"d_nm": d, "cos": c, "rgb": list(ic.color_linear(d, c))} for d, c in lin_pts],
    "gamma": [{"d_nm": d, "cos": c, "rgb": list(ic.color(d, c))} for d, c in gam_pts],
}
with open(os.path.join(os.path.dirname(__file__), "kat_interference.json"), "w") as f:
    json.dump(kat, f, indent=1)
print("wrote kat_interference.json")
