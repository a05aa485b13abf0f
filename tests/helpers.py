"""Shared test helpers: configs, oracle construction, parity metrics."""
from __future__ import annotations

import os

import numpy as np

from oracle import refapi
from oracle.portapi import PortSim
from pd_mg_pin_corrosion_b200.config import Config

CONFIG_DIR = refapi.CONFIG_DIR

# name -> (dim, cfg file, overrides)
CASES = {
    "2d_default": (2, "params.cfg", {}),
    "2d_poiseuille": (2, "params_poiseuille.cfg", {}),
    "3d_default": (3, "params.cfg", {}),
    # small 3D tube (fast everywhere): same physics, shorter/narrower
    "3d_small": (3, "params.cfg", {"R_tube": 60e-6, "R_wire": 20e-6, "L_wire": 60e-6, "L_upstream": 40e-6,
                                   "L_downstream": 40e-6, "grain_size_mean": 20e-6}),
    # off-lattice geometry: radii/lengths that are not multiples of dx
    "2d_offgrid": (2, "params.cfg", {"dx": 4.3e-6, "R_tube": 101e-6, "R_wire": 33e-6, "L_wire": 207e-6,
                                     "L_upstream": 111e-6, "L_downstream": 93e-6}),
    "3d_offgrid": (3, "params.cfg", {"dx": 4.3e-6, "R_tube": 51e-6, "R_wire": 17e-6, "L_wire": 47e-6,
                                     "L_upstream": 31e-6, "L_downstream": 29e-6, "grain_size_mean": 15e-6}),
    # phase change fires within a few hundred explicit steps
    "2d_dissolve": (2, "params.cfg", {"D_grain": 5e-11, "D_gb": 5e-9, "C_thresh": 0.999,
                                      "corrosion_steps_per_check": 50, "flow_max_iters": 300, "T_final": 6e-4,
                                      "output_every_corr": 10}),
}


def load_cfg(case: str, extra: dict | None = None) -> tuple[int, Config, dict]:
    dim, base, ov = CASES[case]
    ov = dict(ov)
    ov["use_implicit"] = 0
    ov.update(extra or {})
    cfg = Config.load(os.path.join(CONFIG_DIR, base), ov, quiet=True)
    return dim, cfg, ov


_REF_CACHE: dict = {}


def make_ref(case: str, extra: dict | None = None, threads: int = 4):
    """The compiled reference (oracle/_ref) if it was built, else the plain-C port.

    Instances of the plain geometry cases are cached per session (the 3D reference CSR
    takes ~20 s to build); their fields are re-initialised on every request."""
    dim, base, ov = CASES[case]
    ov = dict(ov)
    ov.update(extra or {})
    cacheable = not extra and case != "2d_dissolve"
    if cacheable and case in _REF_CACHE:
        r = _REF_CACHE[case]
        if isinstance(r, refapi.RefSim):
            r.lib.ref_fields_init(r.h)
        else:
            r.init_fields(r.is_gb.copy(), r.is_precip.copy())
        return r
    if refapi.have_ref(dim):
        r = refapi.RefSim(dim, base, ov, threads=threads)
    else:
        r = make_port(case, extra, threads)
        r.init_fields()
    if cacheable:
        _REF_CACHE[case] = r
    return r


def make_port(case: str, extra: dict | None = None, threads: int = 4, state_from=None) -> PortSim:
    dim, cfg, _ = load_cfg(case, extra)
    p = PortSim(dim, cfg, threads=threads)
    if state_from is not None:
        p.load_state(state_from)
    return p


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max_i |a_i - b_i| / max_i |b_i|  (SURVEY.md 7.3 field tolerance)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    m = float(np.max(np.abs(b))) if b.size else 0.0
    return d / m if m > 0 else d


def perturbed_state(ref, seed: int = 0):
    """Deterministic non-trivial state written into the reference object: Poiseuille + noise."""
    rng = np.random.default_rng(seed)
    nt = ref.get("node_type") if hasattr(ref, "get") else ref.node_type
    N, dim = nt.size, ref.dim
    inside = (nt != 5)
    rho = ref.get("rho") * (1.0 + 1e-4 * rng.standard_normal(N)) * inside
    vel = (ref.get("vel") * (1.0 + 1e-2 * rng.standard_normal((N, dim)))
           + 1e-3 * rng.standard_normal((N, dim))) * inside[:, None]
    C = np.abs(ref.get("C") + 0.02 * rng.standard_normal(N)) * inside
    ref.set("rho", rho)
    ref.set("vel", vel)
    ref.set("C", C)
    ref.set("rho_new", rho)
    ref.set("vel_new", vel)
    ref.set("C_new", C)
    return rho, vel, C
