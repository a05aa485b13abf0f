"""Re-type the parameter VALUES of the reference's shipped configuration files into configs/ (run in the build
container only; /root/reference is not on the GPU box).  Only `key = value` pairs are taken (the last occurrence of a key, as in the parser), written in this repository's key
order and in shortest round-trip notation; the reference's comments are not copied.  tests/test_config.py checks that
the reference's own parser reads the same values from both files.
usage: python tools/gen_configs.py [--all]    (default: only the files configs/ does not hold yet)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/config"
NOTE = {
    "params.cfg": "80 um Mg-4Ag wire in an SBF flow cell, dx = 5 um, m = 3 (BASELINE configs 1 and 3); as shipped it runs the\n# implicit branch -- the north-star path is the EXPLICIT branch: use_implicit = 0 (the tests / bench override it)",
    "params_fine.cfg": "dx = 2 um resolution of the params.cfg geometry (BASELINE config 4; bench.py's workload with use_implicit = 0)",
    "params_poiseuille.cfg": "Flow-only 2D Poiseuille validation, no wire (BASELINE config 2)",
    "params_amr.cfg": "Two-level AMR production set-up: dx = 2.5 um around the wire, 7.5 um in the far field, R_tube = 425 um; built by\n# pdamr_* (csrc/amr.cu), 2D, implicit branch",
    "params_amr_r2.cfg": "Two-level AMR set-up with refinement ratio 2",
    "params_calibration.cfg": "Calibration run of the corrosion parameters (2D)",
    "params_calibration_v2.cfg": "Calibration run of the corrosion parameters, second parameter set (2D)",
    "params_diagnostic.cfg": "Short diagnostic run",
    "params_fine_calibration.cfg": "Calibration run on the fine lattice",
    "params_implicit_test.cfg": "Implicit-branch test run",
    "params_transport_viz.cfg": "Transport visualisation run",
}

sys.path.insert(0, ROOT)
from dataclasses import fields as dc_fields  # noqa: E402

from pd_mg_pin_corrosion_b200.config import Config  # noqa: E402

ORDER = [f.name for f in dc_fields(Config)]       # this repository's own grouping of the keys
TYPES = {f.name: f.type for f in dc_fields(Config)}


def norm(key: str, text: str) -> str:
    """shortest text that parses to the same value"""
    t = TYPES[key]
    if t in ("int", int):
        return str(int(text))
    if t in ("float", float):
        return repr(float(text))
    return text


for name in sorted(os.listdir(REF)):
    dst = os.path.join(ROOT, "configs", name)
    if os.path.exists(dst) and "--all" not in sys.argv:
        continue
    vals = {}
    for line in open(os.path.join(REF, name)):
        line = line.split("#", 1)[0].strip()
        if "=" not in line:
            continue
        k, v = (s.strip() for s in line.split("=", 1))
        if k and v and k in TYPES:
            vals[k] = v                              # later keys win, as in the parser
    with open(dst, "w") as f:
        f.write(f"# {NOTE.get(name, name)} (parameter values as in the reference's config/{name}).\n")
        f.write("# key = value, read by pd_mg_pin_corrosion_b200/config.py and host/config.cpp like the reference's Config::load;\n")
        f.write("# keys in the order of pd_mg_pin_corrosion_b200/config.py, values in shortest round-trip notation.\n")
        for k in ORDER:
            if k in vals:
                f.write(f"{k} = {norm(k, vals[k])}\n")
    print("wrote", dst, len(vals), "keys")
