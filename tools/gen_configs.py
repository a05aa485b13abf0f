"""Re-type the parameter VALUES of the reference's shipped configuration files into configs/ (run in the build
container only; /root/reference is not on the GPU box).  Only `key = value` pairs are taken, in file order, with the
value text as written (so both parsers read the same digits); the reference's comments are not copied.
usage: python tools/gen_configs.py [--all]    (default: only the files configs/ does not hold yet)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/config"
NOTE = {
    "params_amr_r2.cfg": "Two-level AMR set-up with refinement ratio 2",
    "params_calibration.cfg": "Calibration run of the corrosion parameters (2D)",
    "params_calibration_v2.cfg": "Calibration run of the corrosion parameters, second parameter set (2D)",
    "params_diagnostic.cfg": "Short diagnostic run",
    "params_fine_calibration.cfg": "Calibration run on the fine lattice",
    "params_implicit_test.cfg": "Implicit-branch test run",
    "params_transport_viz.cfg": "Transport visualisation run",
}

for name in sorted(os.listdir(REF)):
    dst = os.path.join(ROOT, "configs", name)
    if os.path.exists(dst) and "--all" not in sys.argv:
        continue
    rows = []
    for line in open(os.path.join(REF, name)):
        line = line.split("#", 1)[0].strip()
        if "=" not in line:
            continue
        k, v = (s.strip() for s in line.split("=", 1))
        if k and v:
            rows.append((k, v))
    with open(dst, "w") as f:
        f.write(f"# {NOTE.get(name, name)} (parameter values as in the reference's config/{name}).\n")
        f.write("# key = value, read by pd_mg_pin_corrosion_b200/config.py and host/config.cpp like the reference's Config::load.\n")
        for k, v in rows:
            f.write(f"{k} = {v}\n")
    print("wrote", dst, len(rows), "keys")
