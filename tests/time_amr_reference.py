"""TEST INFRASTRUCTURE (measurement helper, not a test): the reference's own AMR loop bodies (oracle/_ref) on the host
cores, for the table in profiles/r2_notes.md beside tools/time_amr.py.  usage: python tests/time_amr_reference.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refapi   # noqa: E402

for th in sorted({1, os.cpu_count()}):
    r = refapi.RefSim(2, "params_amr.cfg", {"use_implicit": 0}, threads=th, build=True, fields=True)
    dt = r.ns_compute_dt()
    r.ns_iterate(5, dt)
    t0 = time.perf_counter(); r.ns_iterate(40, dt); c_ns = (time.perf_counter() - t0) / 40
    dtc = r.ard_compute_dt()
    t0 = time.perf_counter(); r.ard_iterate(40, dtc); c_ard = (time.perf_counter() - t0) / 40
    print(f"reference on {th} host thread(s): NS loop body {1e6 * c_ns:.0f} us, ARD loop body {1e6 * c_ard:.0f} us")
