"""Host-side mirror of the reference's `Config` (src/config.h:4-99, src/config.cpp:16-112).

Same `key = value` file format, `#` comments, unknown-key warning, missing-file
warning + defaults, and the same derived quantities (`delta = m_ratio*dx`,
`U_in = Q/(pi R_tube^2)`, `c0 := max(c0, 25 U_in)`).  `to_struct()` produces the POD
`PdConfig` that crosses the C ABI (include/pdgpu.h) *after* compute_derived, so the
mutated `c0`, `delta` and `U_in` are what the device sees.
"""
from __future__ import annotations

import ctypes as C
import math
import re
import sys
from dataclasses import dataclass, fields as dc_fields

PI = 3.14159265358979323846


class PdConfig(C.Structure):
    """ctypes image of `struct PdConfig` in include/pdgpu.h (member order matters)."""

    _fields_ = (
        [(n, C.c_double) for n in (
            "dx", "R_wire", "L_wire", "R_tube", "L_upstream", "L_downstream",
            "rho_f", "mu_f", "gamma_eos", "c0", "eta_density", "Q_flow", "rho_m",
            "D_liquid", "D_grain", "D_gb", "D_precip",
            "C_solid_init", "C_liquid_init", "C_thresh", "C_sat", "alpha_art_diff",
            "corrosion_decay_l", "cfl_factor", "cfl_factor_corr", "flow_conv_tol", "T_final",
            "delta", "U_in")]
        + [(n, C.c_int) for n in (
            "m_ratio", "flow_max_iters", "corrosion_steps_per_check", "output_every_flow",
            "output_every_corr", "channel_flow_corrections", "use_implicit", "reserved")]
    )


@dataclass
class Config:
    # Grid
    dx: float = 5.0e-6
    m_ratio: int = 3
    # Geometry [m]
    R_wire: float = 40.0e-6
    L_wire: float = 400.0e-6
    R_tube: float = 150.0e-6
    L_upstream: float = 80.0e-6
    L_downstream: float = 80.0e-6
    # Fluid
    rho_f: float = 1000.0
    mu_f: float = 1.0e-3
    gamma_eos: float = 7.0
    c0: float = 0.5
    eta_density: float = 0.1
    Q_flow: float = 1.667e-8
    rho_m: float = 1738.0
    # Transport
    D_liquid: float = 1.0e-9
    D_grain: float = 5.0e-11
    D_gb: float = 5.0e-9
    D_precip: float = 5.0e-15
    precip_fraction: float = 0.05
    C_solid_init: float = 1.0
    C_liquid_init: float = 0.0
    C_thresh: float = 0.2
    C_sat: float = 0.9
    alpha_art_diff: float = 0.1
    corrosion_decay_l: float = 0.0
    # Grains
    grain_size_mean: float = 40.0e-6
    grain_size_std: float = 5.0e-6
    gb_width_cells: int = 1
    precip_cluster_cells: int = 0
    # Time stepping
    cfl_factor: float = 0.25
    cfl_factor_corr: float = 0.25
    # Coupling
    flow_max_iters: int = 50000
    flow_conv_tol: float = 5.0e-6
    T_final: float = 32400.0
    corrosion_steps_per_check: int = 200
    output_every_flow: int = 2000
    output_every_corr: int = 100
    output_dir: str = "output"
    # Implicit branch (src/config.h:72-80): matrix-free operator + GMRES on the device (solver.PD_ARD_ImplicitSolver)
    use_implicit: int = 1
    implicit_dt_fraction: float = 0.5
    implicit_dt_max: float = 60.0
    implicit_output_every: int = 10
    diagnostic_every: int = 1
    newton_tol: float = 1.0e-8
    newton_max_iter: int = 20
    channel_flow_corrections: int = 0
    # AMR keys (amr.AmrGrid; the uniform-lattice Grid needs use_amr = 0)
    use_amr: int = 0
    amr_ratio: int = 3
    amr_buffer: float = 50.0e-6
    # Derived
    delta: float = 0.0
    U_in: float = 0.0
    dx_coarse: float = 0.0
    delta_coarse: float = 0.0

    _DERIVED = ("delta", "U_in", "dx_coarse", "delta_coarse")

    @classmethod
    def load(cls, filename: str | None, overrides: dict | None = None, quiet: bool = False) -> "Config":
        """Config::load (src/config.cpp:16-96): sequential scan, later keys win."""
        cfg = cls()
        types = {f.name: f.type for f in dc_fields(cls)}
        lines: list[str] = []
        if filename is not None:
            try:
                with open(filename) as f:
                    lines = f.read().splitlines()
            except OSError:
                print(f"Warning: Cannot open config file '{filename}', using defaults.", file=sys.stderr)
        for k, v in (overrides or {}).items():
            lines.append(f"{k} = {v!r}" if isinstance(v, float) else f"{k} = {v}")
        for line in lines:
            line = line.split("#", 1)[0].strip()
            if not line or "=" not in line:
                continue
            key, val = (s.strip() for s in line.split("=", 1))
            if not key or not val:
                continue
            if key in cls._DERIVED or key not in types:
                print(f"Warning: Unknown config key '{key}'", file=sys.stderr)
                continue
            t = types[key]
            if t in ("int", int):
                m = re.match(r"[+-]?\d+", val)  # std::stoi: leading integer prefix
                if m is None:
                    raise ValueError(f"config key '{key}': cannot parse integer from '{val}'")
                setattr(cfg, key, int(m.group(0)))
            elif t in ("float", float):
                setattr(cfg, key, float(val))
            else:
                setattr(cfg, key, val)
        cfg.compute_derived(quiet=quiet)
        return cfg

    def compute_derived(self, quiet: bool = False) -> None:
        """Config::compute_derived (src/config.cpp:98-112)."""
        self.delta = self.m_ratio * self.dx
        self.dx_coarse = self.amr_ratio * self.dx
        self.delta_coarse = self.m_ratio * self.dx_coarse
        self.U_in = self.Q_flow / (PI * self.R_tube * self.R_tube)
        if self.c0 < 25.0 * self.U_in:
            self.c0 = 25.0 * self.U_in
            if not quiet:
                print(f"NOTE: Increased c0 to {self.c0:.4e} (25x U_in) for stability.")

    def to_struct(self) -> PdConfig:
        s = PdConfig()
        for name, _ in PdConfig._fields_:
            if name == "reserved":
                continue
            setattr(s, name, getattr(self, name))
        return s

    def check_supported(self) -> None:
        if self.use_amr:
            raise ValueError("use_amr = 1: the two-level AMR cloud is built by amr.AmrGrid (2D; explicit and implicit branch), "
                             "not by the uniform-lattice Grid")

    def describe(self, dim: int) -> str:
        """Config::print (src/config.cpp:114-139)."""
        rows = [
            ("DIM", f"{dim}"), ("dx", f"{self.dx:.2e} m"),
            ("delta", f"{self.delta:.2e} m (m={self.m_ratio})"),
            ("R_wire", f"{self.R_wire:.2e} m"), ("L_wire", f"{self.L_wire:.2e} m"),
            ("R_tube", f"{self.R_tube:.2e} m"), ("U_in", f"{self.U_in:.4e} m/s"),
            ("c0", f"{self.c0:.2f} m/s (Mach ~ {self.U_in / self.c0:.4f})"),
            ("D_liquid", f"{self.D_liquid:.2e} m2/s"), ("D_grain", f"{self.D_grain:.2e} m2/s"),
            ("D_gb", f"{self.D_gb:.2e} m2/s"), ("T_final", f"{self.T_final:.1f} s"),
            ("output_dir", self.output_dir),
        ]
        return "=== Configuration ===\n" + "\n".join(f"  {k:<12} = {v}" for k, v in rows)


def horizon_volume(cfg: Config, dim: int) -> tuple[float, float]:
    """(V_H, beta) of PD_NS_Solver::init / PD_ARD_Solver::init (src/pd_ns.cpp:7-16)."""
    if dim == 2:
        return math.pi * cfg.delta ** 2, 4.0 / (math.pi * cfg.delta ** 2)
    return (4.0 / 3.0) * math.pi * cfg.delta ** 3, 12.0 / (math.pi * cfg.delta ** 2)
