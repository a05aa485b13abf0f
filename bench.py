#!/usr/bin/env python
"""bench.py -- PD bond-updates/s of the fused NS + ARD step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libpdgpu.so on B200)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU/OpenMP path

A "step" = one PD-NS loop body (inlet, outlet, wall, solid BCs + bond kernel + wall mirror of
the new buffers, src/pd_ns.cpp:196-205) followed by one explicit PD-ARD loop body (inlet,
outlet, wall-C BCs + bond kernel, src/coupling.cpp:232-238) on the 3D params_fine geometry
(dx = 2 um, 157 x 157 x 707 nodes, BASELINE config 4).  bond-updates per step = CSR row
lengths summed over FLUID rows (NS) + FLUID and SOLID_MG rows (ARD).
For N > 1 the tube is lengthened N-fold (weak scaling: every GPU owns the same z-slab) and
each rank drives one GPU; halos travel over NCCL.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# CPU-baseline protocol: OpenMP threads pinned. libgomp reads this when it is loaded
# (numpy / torch / the oracle library), so it is set before any of them is imported -- but only in the
# one process that times the CPU path: under torchrun every rank would pin its main thread to the same
# core (measured: 5.4 -> 8.6 ms per step at N = 2).
if int(os.environ.get("WORLD_SIZE", "1")) == 1 or ("reference" in sys.argv and os.environ.get("RANK", "0") == "0"):
    os.environ.setdefault("OMP_PROC_BIND", "true")

import numpy as np  # noqa: E402

from pd_mg_pin_corrosion_b200.config import Config  # noqa: E402

METRIC = "pd_bond_updates_per_s_ns_plus_ard_step"
UNIT = "bond-updates/s"


def workload_cfg(n_gpus: int, sample: bool = False, big: bool = False) -> tuple[Config, str]:
    """3D params_fine (+ use_implicit = 0); tube x n_gpus for weak scaling; `sample` = the
    same cross-section with a short tube (bounded CPU-baseline sample); `big` = BASELINE
    configs[4]: the params_fine tube at dx = 1 um (307x307x1407, 1.84e10 CSR-equivalent entries),
    the SAME domain for every GPU count (strong scaling)."""
    ov = {"use_implicit": 0}
    if big:
        ov["dx"] = 1.0e-6
        return (Config.load(os.path.join(ROOT, "configs", "params_fine.cfg"), ov, quiet=True),
                "synthetic 3D tube: params_fine geometry at dx=1um (307x307x1407), strong scaling")
    base = Config.load(os.path.join(ROOT, "configs", "params_fine.cfg"), ov, quiet=True)
    if sample:
        ov.update({"L_wire": 24e-6, "L_upstream": 16e-6, "L_downstream": 16e-6})
        name = "3D params_fine cross-section (dx=2um, 157x157), tube shortened to 35 axial planes"
    else:
        ov.update({"L_wire": base.L_wire * n_gpus, "L_upstream": base.L_upstream * n_gpus,
                   "L_downstream": base.L_downstream * n_gpus})
        name = "3D params_fine (dx=2um, 157x157x707, use_implicit=0)" + (
            f", tube x{n_gpus} (one params_fine slab per GPU)" if n_gpus > 1 else "")
    return Config.load(os.path.join(ROOT, "configs", "params_fine.cfg"), ov, quiet=True), name


# ------------------------------------------------------------------ clocks ------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:   # region shorter than the sampling period: take whatever was seen
            for ts, line in self.rows[-3:]:
                p = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(p[0])); smax = float(p[1])
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------- CPU reference ---------
class _StdoutToStderr:
    """The reference printf()s progress to fd 1; keep stdout clean for the ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference(steps: int, warmup: int, threads: int | None = None) -> dict:
    with _StdoutToStderr():
        return _cpu_reference(steps, warmup, threads)


def _thread_sweep(nproc: int) -> list[int]:
    """{1, nproc/2, nproc-1, nproc}"""
    return sorted({1, max(1, nproc // 2), max(1, nproc - 1), nproc})


def _cpu_reference(steps: int, warmup: int, threads: int | None = None) -> dict:
    """The reference's own OpenMP path (oracle/_ref, else the plain-C port) on the host cores, on the
    bounded sample of the workload: thread sweep {1, nproc/2, nproc-1, nproc} with OMP_PROC_BIND=true,
    best figure reported with its thread count. Test/baseline infrastructure only."""
    from oracle import refapi
    cfg, name = workload_cfg(1, sample=True)
    nproc = os.cpu_count() or 1
    sweep = [threads] if threads else _thread_sweep(nproc)
    ov = {"L_wire": cfg.L_wire, "L_upstream": cfg.L_upstream, "L_downstream": cfg.L_downstream}
    t_build = time.time()
    if refapi.have_ref(3):
        sim = refapi.RefSim(3, "params_fine.cfg", ov, threads=max(sweep))
        kind = "reference"
        nt = sim.get("node_type")
        rowlen = np.diff(sim.get("nbr_offset").astype(np.int64))
        set_threads = sim.lib.ref_set_threads
    else:
        from oracle.portapi import PortSim
        sim = PortSim(3, cfg, threads=max(sweep))
        sim.init_fields()
        kind = "port"
        nt = sim.node_type
        rowlen = np.diff(sim.csr()[0])
        set_threads = sim.L.pdo_set_threads
    t_build = time.time() - t_build
    ns_bonds = int(rowlen[nt == 0].sum())
    ard_bonds = int(rowlen[(nt == 0) | (nt == 1)].sum())
    dt = sim.ns_compute_dt()
    dtc = sim.ard_compute_dt()
    runs = []
    for thr in sweep:
        if set_threads is not None:
            set_threads(int(thr))
        # steps scaled so that every sweep point costs about the same wall time
        k = max(2, int(round(steps * thr / max(sweep))))
        for _ in range(warmup):
            sim.ns_iterate(1, dt)
            sim.ard_iterate(1, dtc)
        t0 = time.perf_counter()
        for _ in range(k):
            sim.ns_iterate(1, dt)
            sim.ard_iterate(1, dtc)
        el = time.perf_counter() - t0
        runs.append({"threads": int(thr), "steps": k, "seconds": el, "value": (ns_bonds + ard_bonds) * k / el})
    best = max(runs, key=lambda r: r["value"])
    sweep_txt = ", ".join(f"{r['threads']} thr: {r['value'] / 1e9:.3f} G/s" for r in runs)
    return {"value": best["value"], "unit": UNIT, "cores": best["threads"], "kind": kind,
            "sample": f"{name}: {ns_bonds + ard_bonds} bond-updates/step, {best['steps']} steps in "
                      f"{best['seconds']:.2f} s at {best['threads']} of {nproc} host threads, OMP_PROC_BIND="
                      f"{os.environ.get('OMP_PROC_BIND', 'unset')}; sweep [{sweep_txt}] "
                      f"(+{t_build:.1f} s grid/CSR build, untimed)",
            "ms_per_step": 1e3 * best["seconds"] / best["steps"], "sweep": runs, "nproc": nproc}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(2, min(args.steps, 50))   # each step = one NS + one ARD loop body on the bounded sample (~0.1 s at 16 threads)
    base = cpu_reference(steps, min(args.warmup, 1))   # thread sweep, best figure
    _, wname = workload_cfg(args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (deterministic geometry from the cfg; Poiseuille initial flow)",
            "config": {"workload": wname, "measured_on": base["sample"]},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "sweep", "nproc")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def slab_parity(S, L_, dist, local: int, rank: int, world: int) -> dict:
    """In-run check of the multi-GPU path (the driver's test box has one GPU): 3D params.cfg (dx = 5 um,
    67x67x287) advanced by 6 NS + 3 ARD loop bodies on `world` z-slabs and, on rank 0, on one GPU; the
    SHA-256 of every rank's owned planes of rho / vel / C must equal that of the same planes of the
    single-GPU run (same per-node arithmetic and order: SURVEY.md 8e)."""
    import hashlib
    L = L_.load()
    cfg = Config.load(os.path.join(ROOT, "configs", "params.cfg"), {"use_implicit": 0}, quiet=True)

    def advance(grid):
        f = S.Fields()
        f.bind(grid)
        L_.check(L.pdgpu_fields_init(grid.ctx, None, None))
        ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
        ns.init(grid, cfg)
        ard.init(grid, cfg)
        dt = ns.compute_dt(f, grid, cfg)
        ns.iterate(f, grid, cfg, 6, dt)
        dtc = ard.compute_dt(f, grid, cfg)
        ard.iterate(f, grid, cfg, 3, dtc)
        return {n: f.get(n) for n in ("rho", "vel", "C")}

    def digest(arrs, a0, a1, P):
        h = hashlib.sha256()
        for n in ("rho", "vel", "C"):
            h.update(np.ascontiguousarray(arrs[n][a0 * P:a1 * P]).tobytes())
        return h.hexdigest()

    g = S.Grid(3, device=local, rank=rank, nranks=world)
    g.build(cfg)
    g.comm_init_torch()
    mine = (g.a0, g.a1, digest(advance(g), g.a0, g.a1, g.plane))
    P = g.plane
    g.close()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    out = None
    if rank == 0:
        one = S.Grid(3, device=local)
        one.build(cfg)
        ref = advance(one)
        one.close()
        bad = [r for r, (a0, a1, d) in enumerate(parts) if digest(ref, a0, a1, P) != d]
        out = {"slab_parity": "bitwise" if not bad else f"MISMATCH on ranks {bad}",
               "slab_parity_case": f"3D params.cfg (dx=5um, 67x67x287), 6 NS + 3 ARD loop bodies, {world} z-slabs vs 1 GPU, "
                                   "sha-256 of the owned planes of rho/vel/C"}
    return out


# -------------------------------------------------------------------- our arm ---------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-array leg (e2e = null)")
    ap.add_argument("--e2e-chunks", type=int, default=32, help="axial chunks of the host-array step pipeline")
    ap.add_argument("--small", action="store_true", help="dx=5um params.cfg 3D (debug)")
    ap.add_argument("--big", action="store_true",
                    help="BASELINE configs[4]: params_fine at dx=1um (132.6 M nodes), same domain at every N (strong scaling)")
    ap.add_argument("--csr", action="store_true",
                    help="also time the materialised-CSR (reference layout, HBM-bound) bond kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    # NCCL / the libraries may print to fd 1: keep stdout for the ONE JSON line
    guard = _StdoutToStderr()
    guard.__enter__()

    import torch
    from pd_mg_pin_corrosion_b200 import lib as L_, solver as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = L_.load()
    cfg, wname = workload_cfg(world, big=args.big)
    if args.small:
        cfg = Config.load(os.path.join(ROOT, "configs", "params.cfg"), {"use_implicit": 0}, quiet=True)
        wname = "3D params.cfg (dx=5um, 67x67x287) [debug]"
    grid = S.Grid(3, device=local, rank=rank, nranks=world)
    grid.build(cfg)
    if world > 1:
        grid.comm_init_torch()
    for opt in os.environ.get("PDGPU_OPTIONS", "").split(","):   # e.g. PDGPU_OPTIONS=debug_no_halo=1,graph=0
        if "=" in opt:
            k, v = opt.split("=")
            grid.set_option(k.strip(), int(v))
    fields = S.Fields()
    fields.bind(grid)
    L_.check(L.pdgpu_fields_init(grid.ctx, None, None))   # grains: none (flags only matter at the wire surface)
    ns, ard = S.PD_NS_Solver(), S.PD_ARD_Solver()
    ns.init(grid, cfg)
    ard.init(grid, cfg)
    dt = ns.compute_dt(fields, grid, cfg)
    dtc = ard.compute_dt(fields, grid, cfg)
    info = grid.info
    bonds_local = info.ns_bonds + info.ard_bonds

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def steps_resident(n):
        # n x { NS loop body ; ARD loop body } on the device, one host synchronisation at the end
        L_.check(L.pdgpu_step_iterate(grid.ctx, n, dt, dtc))

    steps_resident(args.warmup)
    # ---- device-resident throughput (`value`) ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    L.pdgpu_launch_count(grid.ctx, None, 1)
    barrier()
    t0 = time.time()
    L_.check(L.pdgpu_timer_start(grid.ctx))
    steps_resident(args.steps)
    ms = C.c_float()
    L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms)))
    barrier()
    t1 = time.time()
    launches = C.c_longlong()
    L.pdgpu_launch_count(grid.ctx, C.byref(launches), 0)
    clocks = sampler.stop(t0, t1)
    if world > 1:   # per-rank view on stderr (diagnostics only)
        kk = C.c_float(); ka = C.c_float()
        L_.check(L.pdgpu_time_kernel(grid.ctx, 0, 5, C.byref(kk)))
        L_.check(L.pdgpu_time_kernel(grid.ctx, 1, 5, C.byref(ka)))
        print(f"[bench rank {rank}] planes [{info.a0},{info.a1}) ms/step={ms.value / args.steps:.3f} "
              f"ns_kernel={kk.value:.3f} ard_kernel={ka.value:.3f} outlet_nodes={int(info.counts[4])} "
              f"solid={int(info.counts[1])} fluid={int(info.counts[0])}", file=sys.stderr, flush=True)
    t_ms = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(bonds_local)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(t_ms.item())
    bonds_total = float(tot.item())
    value = bonds_total * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-array API (`e2e`): pinned host state -> H2D -> step -> D2H ----
    n_own = (info.a1 - info.a0) * info.plane
    N = info.N_total
    host = {}
    for name, shape in (("rho", (N,)), ("vel", (N, 3)), ("C", (N,))):
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        host[name] = t.numpy()
        L_.check(L.pdgpu_fields_download(grid.ctx, S._FIELD_IDS[name], host[name].ctypes.data_as(C.c_void_p)))
        host["_t_" + name] = t
    n_up = (min(info.a1 + info.reach, cfg_planes(info)) - max(info.a0 - info.reach, 0)) * info.plane
    h2d = n_up * 8 * 5
    d2h = n_own * 8 * 5

    dim = 3
    vel2d = host["vel"].reshape(N, dim)
    used_chunks = C.c_int(1)
    L_.check(L.pdgpu_step_host_chunks(grid.ctx, args.e2e_chunks, C.byref(used_chunks), None, 0))

    def e2e_step():
        # the call a host-resident driver makes: Fields vectors in, loop bodies on the device, Fields out
        L_.check(L.pdgpu_step_host(grid.ctx, dt, dtc, host["rho"].ctypes.data_as(C.c_void_p),
                                   vel2d.ctypes.data_as(C.c_void_p), host["C"].ctypes.data_as(C.c_void_p),
                                   args.e2e_chunks))

    e2e_steps = 0 if args.no_e2e else max(2, min(args.steps, 5))
    if e2e_steps:
        e2e_step()
    barrier()
    L_.check(L.pdgpu_timer_start(grid.ctx))
    for _ in range(e2e_steps):
        e2e_step()
    ms2 = C.c_float()
    L_.check(L.pdgpu_timer_stop(grid.ctx, C.byref(ms2)))
    barrier()
    t2 = torch.tensor([ms2.value], dtype=torch.float64, device="cuda")
    e2e_per_rank = [ms2.value / max(e2e_steps, 1)]
    if dist is not None:
        parts = [torch.zeros_like(t2) for _ in range(world)]
        dist.all_gather(parts, t2)             # per-rank view: host-memory contention between the ranks shows here
        e2e_per_rank = [float(p.item()) / max(e2e_steps, 1) for p in parts]
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = bonds_total * e2e_steps / (float(t2.item()) * 1e-3) if e2e_steps else None

    # ---- roofline of the dominant kernel (PD-NS bond kernel), timed alone with CUDA events ----
    kms = C.c_float()
    L_.check(L.pdgpu_time_kernel(grid.ctx, 0, 10, C.byref(kms)))
    kms_ard = C.c_float()
    L_.check(L.pdgpu_time_kernel(grid.ctx, 1, 10, C.byref(kms_ard)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    # The offset-table kernels do not stream a CSR: their DRAM traffic is the compulsory field traffic
    # (~1.6 GB per launch, a few % of the HBM peak) and the binding resource is the FP64 pipe.  `frac` is
    # therefore the FP64 fraction: executed useful FP64 work over the DFMA peak measured on this device.
    # The reference-layout (CSR) algorithmic-byte view of SURVEY.md 8d is kept beside it as csr_equiv_*.
    alg_bytes = info.ns_bonds * 44 + int(info.counts[0]) * 65
    csr_equiv = alg_bytes / (kms.value * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ns_kernel_traffic.json")))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    except (OSError, ValueError):
        pass
    fp64_peak = C.c_double()
    fp64_peak3 = C.c_double()
    L_.check(L.pdgpu_fp64_peak(grid.ctx, C.byref(fp64_peak)))
    L_.check(L.pdgpu_fp64_peak3(grid.ctx, C.byref(fp64_peak3)))
    # FP64 work of the NS bond kernel: 10 DFMA-class ops per bond + 4 per staged neighbour value
    # (1.75 bonds per staged value) = 12.3 ops per bond-update, 2 flop each
    ops_per_bond = 10.0 + 4.0 / 1.75
    ns_flops = info.ns_bonds * ops_per_bond * 2.0
    achieved_tf = ns_flops / (kms.value * 1e-3) / 1e12
    streaming = grid.info.m == 3
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak.value, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak.value,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "pd-ns bond kernel (k_ns_stream)" if streaming else "pd-ns bond kernel (generic)",
                "kernel_ms": kms.value,
                "peak_source": "pdgpu_fp64_peak: DFMA micro-benchmark on this device in this run (MEASURED_PEAKS.json "
                               "holds HBM and bf16 figures only)",
                "peak_tflops_3_register_sources": fp64_peak3.value,
                "model": f"{ops_per_bond:.1f} useful FP64 ops per bond-update x 2 flop x {info.ns_bonds} bond-updates per launch; "
                         "tensor cores unused (no dense contraction)",
                "csr_equiv_gbs": csr_equiv, "csr_equiv_frac": csr_equiv / peak, "hbm_peak_gbs": peak,
                "hbm_peak_source": peak_src,
                "csr_equiv_model": "reference CSR layout, 44 B/bond-update + 65 B/FLUID node (SURVEY.md 8d): what the "
                                   "same launch would stream from HBM in the reference's data layout; > 1 because the "
                                   "offset-table formulation removes that stream",
                "algorithmic_bytes_per_launch": alg_bytes,
                "ns_bond_updates_per_s": info.ns_bonds / (kms.value * 1e-3),
                "ard_kernel_ms": kms_ard.value,
                "ard_kernel_what": "ARD bond kernels alone (FLUID tiles + SOLID_MG rows); pre-passes and exchanges untimed",
                "ard_bond_updates_per_s": info.ard_bonds / (kms_ard.value * 1e-3)}

    parity = None
    if world > 1:
        parity = slab_parity(S, L_, dist, local, rank, world)

    csr_view = None
    if args.csr and world == 1:
        # reference-layout CSR on device: 44 B per entry (105 GB for 3D params_fine)
        nnz = grid.build_neighbors()
        grid.set_option("ns_kernel", 3)
        grid.set_option("ard_kernel", 3)
        kc = C.c_float(); ka = C.c_float()
        L_.check(L.pdgpu_time_kernel(grid.ctx, 0, 5, C.byref(kc)))
        L_.check(L.pdgpu_time_kernel(grid.ctx, 1, 5, C.byref(ka)))
        csr_bytes = info.ns_bonds * 44 + int(info.counts[0]) * 65
        csr_view = {"nnz": int(nnz), "csr_gbytes": nnz * 44 / 1e9, "ns_kernel_ms": kc.value, "ard_kernel_ms": ka.value,
                    "ns_bond_updates_per_s": info.ns_bonds / (kc.value * 1e-3),
                    "achieved_gbs": csr_bytes / (kc.value * 1e-3) / 1e9, "frac_of_hbm_peak": csr_bytes / (kc.value * 1e-3) / 1e9 / peak,
                    "what": "k_ns_step_csr: one warp per row streaming the reference CSR layout"}
        grid.set_option("ns_kernel", 2)
        grid.set_option("ard_kernel", 1)
        grid.free_neighbors()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = _cpu_reference(40, 2, None)   # ~4 s of host work on the bounded sample (+ ~2 s CSR build)
        cpu = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample", "sweep", "nproc")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.big else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (deterministic geometry from the cfg; Poiseuille initial flow)",
                "config": {"workload": wname, "nodes": int(N), "bond_updates_per_step": int(bonds_total),
                           "parallelism": f"z-slab x{world}", "l2": "per-step working set (>1 GB) exceeds the 126 MB L2",
                           "ns_kernel": "z-streaming tiles" if grid.info.m == 3 else "generic",
                           "grains": "none (is_gb / is_precip = 0: the flags only select the interface diffusivity "
                                     "of wire-surface bonds)",
                           "elided": ["wall_conc_bc (lazy): apply_wall_concentration_bc writes WALL C, which no bond "
                                      "reads (src/pd_ard.cpp:120); device-resident runs evaluate it when WALL C is "
                                      "observed (download / VTI / checkpoint), not every step"]},
                "clocks": clocks, "gpu_launches": int(launches.value),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                        "chunks": int(used_chunks.value),
                        "ms_per_step_per_rank": [round(x, 3) for x in e2e_per_rank] if e2e_steps else None,
                        "what": "pinned host rho/vel/C -> pdgpu_step_host (H2D, NS body + ARD body, D2H "
                                "pipelined over axial chunks), every step"},
                "roofline": roofline}
        if parity is not None:
            line.update(parity)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if csr_view is not None:
            line["csr_path"] = csr_view
    guard.__exit__()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cfg_planes(info) -> int:
    return info.Nz


if __name__ == "__main__":
    main()
