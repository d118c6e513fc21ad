#!/usr/bin/env python
"""bench.py -- NS timestep throughput (Mcell-updates/s) of the B200 path, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # CPU restatement of the reference

A "step" is one full ``navier_stokes_solver`` call (predictor + Poisson RHS + Poisson solve +
correction + pressure update + checks) on the workload named in ``config.workload``:
BASELINE.json configs[1] (3-D periodic Taylor-Green vortex 512^3, fp64, ppp Poisson) at N = 1, and
z-slabs of the same case (512^3 per GPU, weak scaling) at N > 1.

One JSON line is printed by rank 0.  ``value`` = device-resident throughput (fields already in
HBM), ``e2e`` = the same step driven through the public API with HOST (pinned) arrays: every step
pushes u, v, w, p to the device and pulls them back.  ``roofline`` is for the dominant kernel,
timed live with CUDA events on the library's stream; ``cpu_baseline`` times the CPU oracle on a
bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PI = float(np.arccos(-1.0))
# algorithmic bytes per cell per launch
KERNEL_BYTES_PER_CELL = {
    "pred": 104.0, "poisson_rhs": 32.0, "corr": 72.0, "check": 24.0,
    "fft_x_r2c": 16.0, "fft_x_c2r": 16.0, "fft_lines_fwd": 16.0, "fft_lines_inv": 16.0,
    "fft_solve": 16.0, "thomas_fwd": 16.0, "thomas_bwd": 16.0,
    # fused kernels carry the algorithmic bytes of everything they replace (fusing K-DIV / K-CHECK
    # by recompute does not change the denominator)
    "corr_check": 96.0,        # corr 72 + check 24
    "fft_x_r2c_div": 48.0,     # poisson_rhs 32 + fft_x_r2c 16
}
# ncu kernel-name fragment of each timed family, for roofline.traffic (profiles/ncu_traffic.json)
NCU_NAME = {"pred": "k_pred_tma", "corr_check": "k_corr_tma", "corr": "k_corr<", "fft_x_r2c_div": "k_fft_x_r2c",
            "fft_solve": "k_fft_solve", "fft_lines_fwd": "k_fft_lines_r<512, (int)-1", "fft_lines_inv": "k_fft_lines_r<512, 1",
            "fft_x_c2r": "k_fft_x_c2r", "poisson_rhs": "k_rhs", "check": "k_check"}


def load_traffic(kernel, size):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full
    summary of this same command (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(p))
        if t.get("size") != size:
            return None
        frag = NCU_NAME.get(kernel, kernel)
        for name, b in t["kernels"].items():
            if frag in name:
                return b
    except Exception:
        pass
    return None
STEP_BYTES_PER_CELL = 312.0
# two-phase step, 2-D (fen_b200/csrc/multiphase.cu header): algorithmic bytes per cell per launch
MF_KERNEL_BYTES_PER_CELL = {"vof_recon": 64.0, "vof_sweep": 72.0, "mf_props": 48.0, "mf_pred": 112.0,
                            "poisson_rhs": 24.0, "corr": 64.0, "check": 16.0}
# sum over one step: 2 recon + 2 sweeps + props + pred + rhs + Poisson pn (r2c 16, Thomas 2 x 16, c2r 16) + corr + check
MF_STEP_BYTES_PER_CELL = 2 * 64.0 + (64.0 + 80.0) + 48.0 + 112.0 + 24.0 + 64.0 + 64.0 + 16.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.02)       # nvidia-smi itself takes ~40 ms: a few samples even in a 150 ms region

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def init_tgv3d_slab(shape, delta, k0, pinned):
    """Taylor-Green initial condition of config 2 on one z-slab (global plane offset k0), with ghosts
    left to update_ghost_nodes.  Returns host arrays (nx+2, ny+2, nzl+2), Fortran order."""
    import torch
    nx, ny, nzl = shape

    def alloc():
        n = (nx + 2) * (ny + 2) * (nzl + 2)
        t = torch.empty(n, dtype=torch.float64, pin_memory=pinned)
        return t, t.numpy().reshape((nx + 2, ny + 2, nzl + 2), order="F")

    i = np.arange(1, nx + 1, dtype=np.float64)
    j = np.arange(1, ny + 1, dtype=np.float64)
    k = np.arange(1, nzl + 1, dtype=np.float64) + k0
    sx, cxh = np.sin(i * delta), np.cos((i - 0.5) * delta)
    cyh, sy = np.cos((j - 0.5) * delta), np.sin(j * delta)
    czh = np.cos((k - 0.5) * delta)
    keep, out = [], []
    for m in range(4):
        t, a = alloc()
        a[...] = 0.0
        keep.append(t)
        out.append(a)
    u, v, w, p = out
    # plane by plane to keep temporaries small
    uxy = sx[:, None] * cyh[None, :]
    vxy = -cxh[:, None] * sy[None, :]
    pxy = np.cos(2.0 * (i - 0.5) * delta)[:, None] + np.cos(2.0 * (j - 0.5) * delta)[None, :]
    for kk in range(nzl):
        u[1:-1, 1:-1, kk + 1] = uxy * czh[kk]
        v[1:-1, 1:-1, kk + 1] = vxy * czh[kk]
        p[1:-1, 1:-1, kk + 1] = (1.0 / 16.0) * pxy * (np.cos(2.0 * (k[kk] - 0.5) * delta) + 2.0)
    return keep, (u, v, w, p)


def init_channel_slab(shape, gshape, delta, k0, pinned):
    """Channel initial condition of config 3 (oracle/fen_oracle.py: init_channel) on one z-slab: laminar
    Poiseuille profile between the z walls plus a deterministic sinusoidal perturbation."""
    import torch
    nx, ny, nzl = shape
    gnx, gny, gnz = gshape
    amp = 0.05
    Lz = gnz * delta
    i = np.arange(1, nx + 1, dtype=np.float64)
    j = np.arange(1, ny + 1, dtype=np.float64)
    k = np.arange(1, nzl + 1, dtype=np.float64) + k0
    zc = (k - 0.5) * delta
    kx, ky, kz = 2.0 * PI / (gnx * delta), 2.0 * PI / (gny * delta), PI / Lz
    prof = 4.0 * zc * (Lz - zc) / (Lz * Lz)
    keep, out = [], []
    for m in range(4):
        t = torch.empty((nx + 2) * (ny + 2) * (nzl + 2), dtype=torch.float64, pin_memory=pinned)
        a = t.numpy().reshape((nx + 2, ny + 2, nzl + 2), order="F")
        a[...] = 0.0
        keep.append(t)
        out.append(a)
    u, v, w, p = out
    uxy = np.sin(kx * i * delta)[:, None] * np.cos(ky * (j - 0.5) * delta)[None, :]
    vxy = np.cos(kx * (i - 0.5) * delta)[:, None] * np.sin(ky * j * delta)[None, :]
    wxy = np.cos(kx * (i - 0.5) * delta)[:, None] * np.cos(ky * (j - 0.5) * delta)[None, :]
    for kk in range(nzl):
        u[1:-1, 1:-1, kk + 1] = prof[kk] + amp * uxy * np.sin(kz * zc[kk])
        v[1:-1, 1:-1, kk + 1] = amp * vxy * np.sin(kz * zc[kk])
        if int(k[kk]) != gnz:
            w[1:-1, 1:-1, kk + 1] = amp * wxy * np.sin(2.0 * kz * (k[kk] * delta))
    return keep, (u, v, w, p)


def _tgv_ic(n, delta):
    """ghosted Fortran-ordered u, v, p of the 3-D Taylor-Green case (w = 0), periodic ghosts filled analytically"""
    i = np.arange(0, n + 2, dtype=np.float64)
    s_f, c_c = np.sin(i * delta), np.cos((i - 0.5) * delta)
    c2 = np.cos(2.0 * (i - 0.5) * delta)
    u = np.asfortranarray(s_f[:, None, None] * c_c[None, :, None] * c_c[None, None, :])
    v = np.asfortranarray(-(c_c[:, None, None] * s_f[None, :, None] * c_c[None, None, :]))
    p = np.asfortranarray((1.0 / 16.0) * (c2[:, None, None] + c2[None, :, None]) * (c2[None, None, :] + 2.0))
    return u, v, p


def cpu_port(n, steps, warmup, threads=0):
    """Times the CPU restatement of the reference on an n^3 sample of the Taylor-Green workload with all host
    threads.  Preferred: the plain-C/OpenMP restatement (oracle/fen_oracle_c.c); fallback: the numpy/scipy oracle.
    Returns (Mcell-updates/s, seconds per step, cores, description)."""
    cores = os.cpu_count() or 1
    try:
        from oracle import fen_oracle_c as foc
        delta = 2 * PI / float(np.float32(n))
        c = foc.NavierStokesC(n, n, n, delta, 1.0, 0.01, threads=threads or cores)
        u, v, p = _tgv_ic(n, delta)
        c.set(foc.U, u); c.set(foc.V, v); c.set(foc.P, p)
        dt = min(0.25 * delta / 1.0, (1.0 / 6.0) * delta * delta / 0.01)      # set_timestep(U = 1), CFL = 0.25
        c.dt_o = dt
        for s in range(warmup):
            c.navier_stokes_solver(s + 1, dt)
        t0 = time.perf_counter()
        for s in range(steps):
            c.navier_stokes_solver(warmup + s + 1, dt)
        t = time.perf_counter() - t0
        used = c.threads
        assert abs(c.maxdiv) < 1e-10
        c.destroy()
        return n ** 3 * steps / t / 1e6, t / steps, used, "plain-C/OpenMP restatement (oracle/fen_oracle_c.c)"
    except Exception as e:   # no C compiler on the box: the numpy oracle still gives a baseline
        sys.stderr.write("bench.py: C oracle unavailable (%r), timing the numpy oracle\n" % (e,))
    from oracle import fen_oracle as fo
    fo.set_workers(cores)
    G = fo.Grid(n, n, n, 2 * PI, 2 * PI, 2 * PI)
    ns = fo.NavierStokes(G, 1.0, 0.01)
    fo.init_tgv3d(ns)
    ns.CFL = 0.25
    dt = ns.set_timestep(1.0)
    for s in range(warmup):
        ns.navier_stokes_solver(s + 1, dt)
    t0 = time.perf_counter()
    for s in range(steps):
        ns.navier_stokes_solver(warmup + s + 1, dt)
    t = time.perf_counter() - t0
    return n ** 3 * steps / t / 1e6, t / steps, cores, "numpy/scipy oracle (oracle/fen_oracle.py)"


def cpu_port_channel(n, steps, warmup, threads=0):
    """The plain-C/OpenMP restatement (oracle/fen_oracle_c.c, zwalls) on a 2n x 2n x n sample of the channel workload
    (BASELINE configs[2]: walls in z, ppn Poisson).  Returns (Mcell-updates/s, seconds per step, threads)."""
    from oracle import fen_oracle_c as foc
    cores = os.cpu_count() or 1
    nx, ny, nz = 2 * n, 2 * n, n
    delta = 2.0 / float(np.float32(nx))
    c = foc.NavierStokesC(nx, ny, nz, delta, 1.0, 1.0e-3, threads=threads or cores, zwalls=True)
    _, (u, v, w, p) = init_channel_slab((nx, ny, nz), (nx, ny, nz), delta, 0, pinned=False)
    # ghost cells of the initial condition: periodic in x / y, no-slip at the z walls (what update_ghost_nodes gives)
    for a, kind in ((u, 1), (v, 1), (w, 2), (p, 0)):
        a[0, :, :] = a[nx, :, :]; a[nx + 1, :, :] = a[1, :, :]
        a[:, 0, :] = a[:, ny, :]; a[:, ny + 1, :] = a[:, 1, :]
        if kind == 0:
            a[:, :, 0] = a[:, :, 1]; a[:, :, nz + 1] = a[:, :, nz]
        elif kind == 1:
            a[:, :, 0] = -a[:, :, 1]; a[:, :, nz + 1] = -a[:, :, nz]
        else:
            a[:, :, 0] = 0.0; a[:, :, nz] = 0.0; a[:, :, nz + 1] = 0.0
    c.set(foc.U, u); c.set(foc.V, v); c.set(foc.W, w); c.set(foc.P, p)
    c.g = [1.0, 0.0, 0.0]
    dt = min(0.25 * delta / 1.5, (1.0 / 6.0) * delta * delta / 1.0e-3)       # set_timestep(U = 1.5), CFL = 0.25
    c.dt_o = dt
    for s in range(warmup):
        c.navier_stokes_solver(s + 1, dt)
    t0 = time.perf_counter()
    for s in range(steps):
        c.navier_stokes_solver(warmup + s + 1, dt)
    t = time.perf_counter() - t0
    used = c.threads
    assert abs(c.maxdiv) < 1e-9
    c.destroy()
    return nx * ny * nz * steps / t / 1e6, t / steps, used


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference (oracle/), all host threads, on a bounded sample of
    the same workload (the reference's own Fortran/MPI/FFTW build does not exist in this image)."""
    if rank != 0:
        return
    if args.case == "wave2d":
        val, sec = cpu_port_wave2d(512, 1024, args.steps, args.warmup)
        sample = ("512 x 1024 sample of the 2048 x 4096 two-phase wave, %d steps, plain-C/OpenMP restatement "
                  "(oracle/fen_oracle_mf_c.c)" % args.steps)
        print(json.dumps({
            "impl": "reference", "metric": "two-phase NS timestep Mcell-updates/s", "value": val,
            "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "2-D two-phase gravity wave fp64 (CPU arm: bounded 512 x 1024 sample)",
                       "grid": [512, 1024, 1]},
            "cpu_baseline": {"value": val, "unit": "Mcell-updates/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    if args.case == "channel":
        n = min(args.cpu_size, 128)
        val, sec, cores = cpu_port_channel(n, args.steps, args.warmup)
        sample = ("%dx%dx%d sample of the 1024x1024x512 channel, %d steps, plain-C/OpenMP restatement "
                  "(oracle/fen_oracle_c.c)" % (2 * n, 2 * n, n, args.steps))
        print(json.dumps({
            "impl": "reference", "metric": "NS timestep Mcell-updates/s", "value": val, "unit": "Mcell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "channel fp64, walls in z, ppn Poisson (CPU arm: bounded %dx%dx%d sample)"
                                   % (2 * n, 2 * n, n), "grid": [2 * n, 2 * n, n], "nu": 1.0e-3, "CFL": 0.25},
            "cpu_baseline": {"value": val, "unit": "Mcell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    n = args.cpu_size
    val, sec, cores, what = cpu_port(n, args.steps, args.warmup)
    sample = ("the whole %d^3 Taylor-Green case, %d steps, %s" % (n, args.steps, what)) if n == args.size else \
        ("%d^3 sample of the %d^3 Taylor-Green case, %d steps, %s" % (n, args.size, args.steps, what))
    line = {
        "impl": "reference", "metric": "NS timestep Mcell-updates/s", "value": val, "unit": "Mcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the GPU arm's own configuration when n == --size (default: 512^3, the whole workload, not a sample)
        "config": {"workload": ("3D periodic Taylor-Green vortex %d^3 per GPU fp64, ppp FFT Poisson (BASELINE "
                                "configs[1])" % n) if n == args.size else
                               ("3D periodic Taylor-Green vortex %d^3 fp64, ppp Poisson (CPU arm: bounded %d^3 sample)"
                                % (args.size, n)), "grid": [n, n, n], "nu": 0.01, "CFL": 0.25,
                   "decomposition": "host threads (OpenMP), one process"},
        "cpu_baseline": {"value": val, "unit": "Mcell-updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "restatement of the reference's CPU path (oracle/), not the gfortran/FFTW/2decomp "
                                 "binary (no Fortran toolchain in this image)"},
        "e2e": {"value": val, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(budget_s=20.0, n=512):
    # one probing step decides how many steps fit the budget
    _, sec, _, _ = cpu_port(n, 1, 1)
    steps = int(max(3, min(40, budget_s / max(sec, 1e-3))))
    val, sec, cores, what = cpu_port(n, steps, 1)
    return {"value": val, "unit": "Mcell-updates/s", "cores": cores, "kind": "port",
            "sample": "%d steps of the same Taylor-Green case at %d^3 (1 warm-up), %s" % (steps, n, what)}


def cpu_port_wave2d(nx, ny, steps, warmup=1):
    """The plain-C/OpenMP restatement of the reference's two-phase step (oracle/fen_oracle_mf_c.c, all host threads) on
    an nx x ny sample of the wave2d workload; the numpy restatement only builds the initial state.  Returns
    (Mcell-updates/s, seconds per step)."""
    import math
    from oracle import fen_oracle as fo
    from oracle import fen_oracle_mf as mf
    from oracle import fen_oracle_mf_c as mfc
    fo.set_workers(os.cpu_count() or 1)
    Lx, Ly = 1.0, float(ny) / nx
    G = fo.Grid(nx, ny, 1, Lx, Ly, Lx / nx, bc=["Periodic", "Periodic", "Wall", "Wall"])
    g = 9.80665
    mu0 = 1000.0 * Lx * math.sqrt(g * Lx) / 1.0e4
    ns = mf.MultiphaseNavierStokes(G, 1000.0, 1000.0 / 850.0, mu0, mu0 * 1.9e-2, 0.07,
                                   distance=lambda x, y: y - 0.02 * np.cos(2.0 * PI * x / Lx) - Ly / 2.0)
    ns.g[1] = -g
    dt = 0.1 * ns.set_timestep(1.0)
    c = mfc.MultiphaseC.from_oracle(ns, threads=os.cpu_count() or 1)
    for s in range(warmup):
        c.navier_stokes_solver(s + 1, dt)
    t0 = time.perf_counter()
    for s in range(steps):
        c.navier_stokes_solver(warmup + s + 1, dt)
    t = time.perf_counter() - t0
    assert abs(c.maxdiv) < 1e-6
    c.destroy()
    return nx * ny * steps / t / 1e6, t / steps


def cpu_baseline_wave2d(budget_s=15.0):
    nx, ny = 512, 1024
    _, sec = cpu_port_wave2d(nx, ny, 1, 1)
    steps = int(max(2, min(200, budget_s / max(sec, 1e-3))))
    val, sec = cpu_port_wave2d(nx, ny, steps, 1)
    return {"value": val, "unit": "Mcell-updates/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d steps of the same two-phase wave at %d x %d (1 warm-up), plain-C/OpenMP restatement of the "
                      "reference's -DMF step (oracle/fen_oracle_mf_c.c)" % (steps, nx, ny)}


def run_wave2d(args, local_rank, headline=True):
    """--case wave2d: the two-phase step (MTHINC VoF + one-fluid NS with pressure splitting) on a
    2-D gravity wave between fluids of density ratio 850 -- the 2-D analogue of BASELINE configs[4], which the
    reference cannot express in 3-D (its VoF and variable-viscosity terms have no z part).  One GPU."""
    import math
    import torch
    import fen_b200 as fb
    nx, ny = 2048, 4096         # x is an FFT direction (<= 2048 points); y is solved by the Thomas algorithm
    if args.grid:
        nx, ny = [int(q) for q in args.grid.split(",")][:2]
    Lx, Ly = 1.0, float(ny) / nx
    G = fb.grid().setup(nx, ny, 1, Lx, Ly, Lx / nx, bc=["Periodic", "Periodic", "Wall", "Wall"], device=local_rank)
    ns = fb.MultiphaseSolver(G)
    g = 9.80665
    ns.rho_0 = 1000.0
    ns.rho_1 = 1000.0 / 850.0
    ns.mu_0 = 1000.0 * Lx * math.sqrt(g * Lx) / 1.0e4
    ns.mu_1 = ns.mu_0 * 1.9e-2
    ns.sigma = 0.07
    ns.g = [0.0, -g, 0.0]
    ns.init_solver(None)        # synthetic initial vof pushed below (a Python distance callback per cell is no bench)
    d = G.delta
    x = ((np.arange(0, nx + 2) - 0.5) * d)[:, None]
    y = ((np.arange(0, ny + 2) - 0.5) * d)[None, :]
    a, wn = 0.02, 2.0 * PI / Lx
    ns.vof.f[:, :, 1] = 0.5 * (1.0 + np.tanh(1.0 * (y - a * np.cos(wn * x) - Ly / 2.0) / d))
    ns.vof.push()
    ns.vof.update_ghost_nodes()
    ns.update_material_properties()
    om = math.sqrt(g * wn)
    f = ns.vof.f[:, :, 1]
    xs, ys = x + 0.5 * d, y - Ly / 2.0
    ns.v.x.f[:, :, 1] = ((1.0 - f) * np.exp(np.minimum(wn * ys, 0.0)) - f * np.exp(np.minimum(-wn * ys, 0.0))) * a * om * np.cos(wn * xs)
    xs, ys = x, y + 0.5 * d - Ly / 2.0
    ns.v.y.f[:, :, 1] = ((1.0 - f) * np.exp(np.minimum(wn * ys, 0.0)) + f * np.exp(np.minimum(-wn * ys, 0.0))) * a * om * np.sin(wn * xs)
    ns.v.push()
    ns.v.update_ghost_nodes()
    dt = 0.1 * ns.set_timestep(1.0)
    G.synchronize()
    stream = torch.cuda.ExternalStream(G.lib.fen_gpu_stream(G.ctx))
    step_no = [0]

    def dev_step(_):
        step_no[0] += 1
        ns.navier_stokes_solver(step_no[0], dt)

    def timed(fn, k):
        G.synchronize(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for q in range(k):
            fn(q)
        G.synchronize()          # asynchronous pulls: the region ends when they are on the host
        e1.record(stream)
        G.synchronize(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for q in range(max(args.warmup, 6)):      # x_first alternates: both step graphs are captured during warm-up
        dev_step(q)
    l0 = ns.launch_count()
    with ClockSampler(local_rank) as cs:
        ms = timed(dev_step, args.steps)
    launches = ns.launch_count() - l0
    maxdiv, maxcfl = ns.status()
    i1, i2 = ns.check_vof_integral()
    ncell = nx * ny
    ns.profile(True)
    nprof = min(args.steps, 4)
    for q in range(nprof):
        dev_step(q)
    prof = ns.profile_read()
    ns.profile(False)
    peak, peak_src = load_peaks()
    kernels = []
    for name, (tms, cnt) in prof.items():
        if cnt == 0:
            continue
        per = tms / cnt
        bpc = MF_KERNEL_BYTES_PER_CELL.get(name)
        gbs = (bpc * ncell / (per * 1e-3) / 1e9) if bpc else None
        kernels.append({"kernel": name, "launches_per_step": cnt / nprof, "ms_per_launch": per,
                        "ms_per_step": tms / nprof, "alg_bytes_per_cell": bpc, "achieved_GBs": gbs,
                        "frac": (gbs / peak) if gbs else None})
    kernels.sort(key=lambda k: -k["ms_per_step"])
    dom = next((k for k in kernels if k["achieved_GBs"]), None)
    per_step = ms / args.steps
    step_gbs = MF_STEP_BYTES_PER_CELL * ncell / (per_step * 1e-3) / 1e9
    e2e = None
    if headline and not args.no_e2e:
        # host side of a 2-D driver: interior arrays (nx, ny, 1) -- a 2-D FEN array with ghosts carries two unused z
        # planes -- in pinned memory, attached to the solver's device fields
        from fen_b200.api import VX, VY, P as PID, VOF
        fields, keep = [], []
        for fid, loc in ((VX, "x"), (VY, "y"), (PID, "c"), (VOF, "c")):
            h = fb.scalar(G, 0, loc, field_id=fid)
            t = torch.empty(h.f.size, dtype=torch.float64, pin_memory=True)
            h.f = t.numpy().reshape(h.f.shape, order="F")
            keep.append(t)
            h.pull()
            fields.append(h)
        nbytes = sum(q.f.nbytes for q in fields)

        def e2e_step(_):
            for q in fields:
                q.push()
            ns.v.update_ghost_nodes()      # what a driver does after writing interiors (viscous_decay.f90:129)
            ns.p.update_ghost_nodes()
            ns.vof.update_ghost_nodes()
            dev_step(0)
            ns.status()
            for q in fields:
                q.pull_async()
        e2e_step(0)
        ms_e = timed(e2e_step, args.e2e_steps)
        e2e = {"value": ncell * args.e2e_steps / (ms_e * 1e-3) / 1e6, "unit": "Mcell-updates/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 16, "ms_per_step": ms_e / args.e2e_steps,
               "what": "push the interiors of u, v, p, vof from pinned host arrays + ghost updates + two-phase step + "
                       "status + asynchronous pull of the four interiors, every step"}
    if getattr(args, "affinity_before", None):
        os.sched_setaffinity(0, args.affinity_before)       # the CPU baseline sees every core
    line = {
        "metric": "two-phase NS timestep Mcell-updates/s", "value": ncell * args.steps / (ms * 1e-3) / 1e6,
        "unit": "Mcell-updates/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 6),
        "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "2-D two-phase gravity wave %dx%d fp64 (MTHINC VoF, density ratio 850, constant-"
                               "coefficient pressure splitting, pn Poisson): the 2-D analogue of BASELINE configs[4]"
                               % (nx, ny), "grid": [nx, ny, 1], "dt": dt,
                   "l2": "working set (%d fields x %.0f MB) exceeds the 126 MB L2" % (27, ncell * 8 / 1e6)},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": cs.summary(),
        "roofline": {"bound": "hbm", "kernel": dom["kernel"] if dom else None,
                     "achieved": dom["achieved_GBs"] if dom else None, "peak": peak, "unit": "GB/s",
                     "frac": dom["frac"] if dom else None, "traffic": None, "peak_source": peak_src,
                     "alg_bytes_per_launch": dom["alg_bytes_per_cell"] * ncell if dom else None,
                     "step_alg_bytes_per_cell": MF_STEP_BYTES_PER_CELL, "step_achieved": step_gbs,
                     "step_frac": step_gbs / peak},
        "cpu_baseline": None if (args.no_cpu_baseline or not headline) else cpu_baseline_wave2d(), "kernels": kernels,
        "check": {"maxdiv": maxdiv, "maxCFL": maxcfl, "phase_integrals": [i1, i2]}}
    G.destroy()
    return line


def bind_to_gpu_cpus(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU (the socket the GPU's PCIe root hangs off), so
    that the pinned host arrays of the e2e loop are allocated on that NUMA node -- the usual placement rule for
    host<->device copies (NCCL does the same for its proxy threads).  Returns (previous mask, description); the caller
    restores the mask before the CPU baseline, which must see every core.  FEN_BENCH_NO_AFFINITY=1 leaves it alone."""
    try:
        old = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return None, "unchanged (no sched_getaffinity)"
    if os.environ.get("FEN_BENCH_NO_AFFINITY"):
        return old, "unchanged (FEN_BENCH_NO_AFFINITY)"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(ncpu, max(old) + 1) + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        local = cpus & set(old)
        if not local:
            return old, "unchanged (NVML reports no usable local CPUs)"
        if local == set(old):
            return old, "unchanged (all %d CPUs are local to the GPU)" % len(old)
        if len(local) < 4 or 4 * len(local) < len(old):      # a sliver of the allowed CPUs (cgroup cpuset): not worth it
            return old, "unchanged (only %d of the %d allowed CPUs are local to the GPU)" % (len(local), len(old))
        os.sched_setaffinity(0, local)
        return old, "GPU-local CPUs (%d of %d) while the host arrays are allocated and copied" % (len(local), len(old))
    except Exception as exc:            # a placement hint, never a reason to lose the bench line
        return old, "unchanged (%s)" % repr(exc)[:80]


def nccl_alltoall_reference(nx, ny, nz, world, x_periodic=True, iters=5):
    """What one y<->z transpose of the Poisson solver costs when it is a separate library collective: NCCL
    ``all_to_all_single`` over this rank's spectral slab in ``world`` equal blocks (the plain replacement of 2decomp's
    transpose_y_to_z, src/poisson.f90:982,1015).  The pack / unpack passes such an implementation also needs are NOT
    timed, so this is a lower bound of an NCCL-based transpose; the product's transposes are the epilogues of its own
    FFT / Thomas kernels (poisson.cu), reported beside it.  Max over ranks, CUDA events on torch's current stream
    (the stream NCCL is enqueued from)."""
    import torch
    import torch.distributed as dist
    from fen_b200 import decomp
    sent = decomp.alltoall_bytes_per_gpu(nx, ny, nz, world, x_periodic)
    n = sent // (world - 1) // 8 * world          # doubles in the whole slab, a multiple of world
    src = torch.zeros(n, dtype=torch.float64, device="cuda")
    dst = torch.empty_like(src)
    for _ in range(2):
        dist.all_to_all_single(dst, src)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dist.all_to_all_single(dst, src)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    del src, dst
    return {"ms": ms, "bus_GBs": sent / (ms * 1e-3) / 1e9, "frac": sent / (ms * 1e-3) / 1e9 / 900.0,
            "what": "torch.distributed all_to_all_single (NCCL) of the same slab, %d iterations, no pack/unpack" % iters}


def _ctx():
    return {"rank": int(os.environ.get("RANK", "0")), "world": int(os.environ.get("WORLD_SIZE", "1")),
            "local_rank": int(os.environ.get("LOCAL_RANK", "0"))}


def _profile_read(G):
    """per-kernel event times of the profiled launches: {name: (ms, launches)} (fen_gpu_profile_read)"""
    import ctypes as C
    from fen_b200.api import check
    n = 64
    names = ((C.c_char * 32) * n)()
    ms = (C.c_double * n)()
    cnt = (C.c_int * n)()
    nout = C.c_int()
    check(G.lib.fen_gpu_profile_read(G.ctx, n, names, ms, cnt, C.byref(nout)))
    return {names[i].value.decode(): (ms[i], cnt[i]) for i in range(nout.value)}


def _make_timers(G, torch, dist, world, stream):
    def barrier():
        G.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        G.synchronize()
        torch.cuda.synchronize()

    def timed(fn, k, drain=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(k):
            fn(s)
        if drain:
            G.synchronize()      # asynchronous pulls run on their own stream: the region ends when they are on the host
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    return barrier, timed


def weak_dims(n, world):
    """512^3 cells per GPU, the shapes of BASELINE configs[3] / SURVEY.md 8(d) config 4: 512^3 (1), 1024x512x512 (2),
    1024x1024x512 (4), 1024^3 (8) -- x, then y, then z doubles; z-slabs"""
    dims = [n, n, n]
    for q in range(max(world, 1).bit_length() - 1):
        dims[q % 3] *= 2
    if dims[0] * dims[1] * dims[2] != n ** 3 * world:
        raise SystemExit("bench.py: --gpus must be a power of two (got %d)" % world)
    return dims


def run_poisson_case(args, cx, steps):
    """BASELINE configs[3]: Poisson-only (ppp), 512^3 per GPU; rhs = the reference's analytic test rhs
    (test/small_test/poisson/convergence_rate/convergence_rate.f90:174-176).  Returns the JSON line as a dict."""
    import torch
    import torch.distributed as dist
    import fen_b200 as fb
    rank, world, local_rank = cx["rank"], cx["world"], cx["local_rank"]
    n = args.size
    nx, ny, nz = weak_dims(n, world)
    if args.grid:
        nx, ny, nz = [int(v) for v in args.grid.split(",")]
        n = min(nx, ny, nz)
    L = 2 * PI
    G = fb.grid().setup(nx, ny, nz, L * nx / n, L * ny / n, L * nz / n, pcol=world, rank=rank, device=local_rank)
    if world > 1:
        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        G.connect(all_gather)
    phi = fb.scalar(G, 1)
    ps = fb.PoissonSolver(phi)
    stream = torch.cuda.ExternalStream(G.lib.fen_gpu_stream(G.ctx))
    barrier, timed = _make_timers(G, torch, dist, world, stream)
    # rhs = lap(f) for f = sin(kx x) cos(ky y) sin(kz z), one wave per box side
    d = G.delta
    x = (np.arange(1, nx + 1) - 0.5) * d
    y = (np.arange(1, ny + 1) - 0.5) * d
    z = (np.arange(G.lo[2], G.hi[2] + 1) - 0.5) * d
    kxw, kyw, kz = float(n) / nx, float(n) / ny, float(n) / nz
    sxy = np.sin(kxw * x)[:, None] * np.cos(kyw * y)[None, :]
    for kk in range(G.nloc[2]):
        phi.f[1:-1, 1:-1, kk + 1] = -(kxw * kxw + kyw * kyw + kz * kz) * sxy * np.sin(kz * z[kk])
    rhs_keep = phi.f.copy()
    phi.push()
    G.synchronize()
    lib, ctx = G.lib, G.ctx

    def solve(_):
        fb.api.check(lib.fen_gpu_solve_poisson(ctx, phi.id))
    for s in range(max(args.warmup, 3)):
        solve(s)
    l0 = lib.fen_gpu_launch_count(ctx)
    with ClockSampler(local_rank) as cs:
        ms = timed(solve, steps)
    launches = lib.fen_gpu_launch_count(ctx) - l0
    barrier()
    nprof = min(steps, 5)
    fb.api.check(lib.fen_gpu_profile_enable(ctx, 1))
    for s in range(nprof):
        solve(s)
    prof = _profile_read(G)
    fb.api.check(lib.fen_gpu_profile_enable(ctx, 0))
    peak, peak_src = load_peaks()
    ncell_loc = G.nloc[0] * G.nloc[1] * G.nloc[2]
    per = ms / steps
    kernels = [{"kernel": k, "ms_per_solve": t / nprof} for k, (t, cnt) in prof.items() if cnt]
    kernels.sort(key=lambda q: -q["ms_per_solve"])
    ach = 80.0 * ncell_loc / (per * 1e-3) / 1e9
    nvlink = nvlink_figures(args, {q["kernel"]: q["ms_per_solve"] for q in kernels}, nx, ny, nz, world, nccl=False)
    # correctness of what was timed: one solve of the analytic rhs against the analytic solution
    phi.f[...] = rhs_keep
    phi.push(); solve(0); phi.pull()
    sol = np.empty_like(phi.f[1:-1, 1:-1, 1:-1])
    for kk in range(G.nloc[2]):
        sol[:, :, kk] = sxy * np.sin(kz * z[kk])
    err = float(np.abs(phi.f[1:-1, 1:-1, 1:-1] - sol).max())
    line = {"metric": "Poisson solve ms", "value": per, "unit": "ms", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Poisson-only ppp %d^3 per GPU (BASELINE configs[3])" % n, "grid": [nx, ny, nz],
                       "decomposition": "z-slabs x%d" % world},
            "gpu_launches": int(launches), "clocks": cs.summary(),
            "roofline": {"bound": "hbm", "kernel": "poisson solve (5 passes)", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "alg_bytes_per_launch": 80.0 * ncell_loc},
            "kernels": kernels, "nvlink": nvlink,
            "check": {"max_error_vs_analytic": err, "second_order_bound": 4.0 * d * d}}
    barrier()
    G.destroy()
    return line


def nvlink_figures(args, tk, nx, ny, nz, world, nccl=True):
    """transposes over NVLink (N > 1): bytes each GPU stores into its peers / time of the kernels that do it"""
    if world <= 1:
        return None
    from fen_b200 import decomp
    sent = decomp.alltoall_bytes_per_gpu(nx, ny, nz, world)
    fwd = tk.get("fft_lines_fwd_a2a", 0.0) + tk.get("a2a_fwd_sync", 0.0)
    bwd = tk.get("fft_solve_a2a", 0.0) + tk.get("a2a_scatter", 0.0) + tk.get("a2a_bwd_sync", 0.0)
    nv = {"a2a_bytes_sent_per_gpu": sent, "peak_GBs_per_dir": 900.0, "measured_peer_copy_GBs": 770.0,
          "fwd_ms": fwd, "fwd_bus_GBs": sent / (fwd * 1e-3) / 1e9 if fwd else None,
          "bwd_ms": bwd, "bwd_bus_GBs": sent / (bwd * 1e-3) / 1e9 if bwd else None,
          "note": "y<->z transposes = epilogues of the y-FFT / z-solve kernels: the transformed tile goes from shared "
                  "memory to the owning ranks as one bulk store (cp.async.bulk) of blk x 128 bytes per destination; "
                  "time = that kernel + the flag handshake on this rank, measured in the profiled steps, where the solve "
                  "runs in ONE piece on one stream, so the figure is a lower bound of the link rate (it includes the "
                  "transform itself); in the timed region the solve runs in pieces that overlap with the x pass / y "
                  "inverse of their neighbours (extra.poisson_only.value is the overlapped solve time); "
                  "measured_peer_copy = B200_PROFILING.md's 770 GB/s"}
    for kname in ("fwd", "bwd"):
        v = nv[kname + "_bus_GBs"]
        nv[kname + "_frac"] = v / 900.0 if v else None
    if nccl and not args.no_nccl_baseline:
        try:
            nv["nccl_alltoall"] = nccl_alltoall_reference(nx, ny, nz, world)
        except Exception as exc:      # a reported baseline: never lose the bench line over it
            nv["nccl_alltoall"] = {"unavailable": repr(exc)[:200]}
    return nv


def cufft_crosscheck(n):
    """Performance cross-check of the hand-written Poisson passes against cuFFT (north_star: "cuFFT used only as a
    correctness and performance cross-check"): the time of torch.fft.rfftn + a spectral divide + torch.fft.irfftn on an
    n^3 fp64 array -- what the ppp solve does -- timed with CUDA events after warm-up.  Library code, never on the
    product path; the correctness half is tests/test_gpu_parity.py::test_poisson_ppp_matches_cufft."""
    import torch
    x = torch.randn((n, n, n), device="cuda", dtype=torch.float64)
    lam = torch.rand((n, n, n // 2 + 1), device="cuda", dtype=torch.float64) + 1.0

    def solve():
        return torch.fft.irfftn(torch.fft.rfftn(x) / lam, s=(n, n, n))
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 10
    e0.record()
    for _ in range(k):
        solve()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    del x, lam
    torch.cuda.empty_cache()
    return {"cufft_rfftn_divide_irfftn_ms": ms, "grid": [n, n, n], "dtype": "f64",
            "what": "torch.fft (cuFFT) 3-D r2c + elementwise divide + c2r of the same size, 10 repetitions; compare with "
                    "extra.poisson_only.value (the hand-written solve, which also carries the real-field ghost writes)"}


def parity_vs_oracle(args, cx, ny, nz):
    """N > 1: one navier_stokes_solver step on a grid with the SAME y / z line lengths and the same slab split as the
    timed one (so the same transpose kernels, block sizes and per-rank line counts), x shrunk to 16 cells so that the
    plain-C restatement of the reference (oracle/fen_oracle_c.c) can step the whole grid on rank 0 in a second.
    Returns {"rel_l2_vs_oracle": {...}} on rank 0.  The oracle is the checker here, never the thing timed."""
    import torch.distributed as dist
    import fen_b200 as fb
    rank, world, local_rank = cx["rank"], cx["world"], cx["local_rank"]
    nx = 16
    d = 2 * PI / float(np.float32(ny))
    G = fb.grid().setup(nx, ny, nz, nx * d, ny * d, nz * d, pcol=world, rank=rank, device=local_rank)

    def all_gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    G.connect(all_gather)
    ns = fb.Solver(G, 1.0, 0.01).init_solver()
    ns.CFL = 0.25
    dt = ns.set_timestep(1.0)
    ax, ay, az = 2 * PI / (nx * d), 2 * PI / (ny * d), 2 * PI / (nz * d)

    def fields(k):          # k: global plane indices (ghosts included), analytic = periodic
        i = np.arange(0, nx + 2, dtype=np.float64)[:, None, None]
        j = np.arange(0, ny + 2, dtype=np.float64)[None, :, None]
        k = np.asarray(k, dtype=np.float64)[None, None, :]
        xf, xc, yf, yc, zf, zc = i * d, (i - 0.5) * d, j * d, (j - 0.5) * d, k * d, (k - 0.5) * d
        u = np.sin(ax * xf) * np.cos(ay * yc) * np.cos(az * zc)
        v = -np.cos(ax * xc) * np.sin(ay * yf) * np.cos(az * zc)
        w = 0.5 * np.cos(ax * xc) * np.cos(ay * yc) * np.sin(az * zf)
        p = (1.0 / 16.0) * (np.cos(2 * ax * xc) + np.cos(2 * ay * yc)) * (np.cos(2 * az * zc) + 2.0)
        return [np.asfortranarray(q) for q in (u, v, w, p)]
    mine = fields(np.arange(G.lo[2] - 1, G.hi[2] + 2))
    for a, q in zip((ns.v.x, ns.v.y, ns.v.z, ns.p), mine):
        a.f[...] = q
        a.push()
    G.synchronize()
    dist.barrier()
    ns.navier_stokes_solver(1, dt)
    maxdiv, _ = ns.status()
    ns.v.pull(); ns.p.pull()
    slabs = [a.f[1:-1, 1:-1, 1:-1].copy() for a in (ns.v.x, ns.v.y, ns.v.z, ns.p)]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(slabs, gathered, dst=0)
    out = None
    if rank == 0:
        from oracle import fen_oracle_c as foc
        co = foc.NavierStokesC(nx, ny, nz, d, 1.0, 0.01)
        co.dt_o = dt
        for fid, q in zip((foc.U, foc.V, foc.W, foc.P), fields(np.arange(0, nz + 2))):
            co.set(fid, q)
        co.navier_stokes_solver(1, dt)
        errs = {}
        for m, (name, fid) in enumerate((("u", foc.U), ("v", foc.V), ("w", foc.W), ("p", foc.P))):
            ref = co.get(fid)[1:-1, 1:-1, 1:-1]
            got = np.concatenate([g_[m] for g_ in gathered], axis=2)
            errs[name] = float(np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel()))
        co.destroy()
        out = {"rel_l2_vs_oracle": errs, "grid": [nx, ny, nz], "steps": 1, "maxdiv": maxdiv,
               "what": "one step on %dx%dx%d (the timed grid's y / z line lengths and slab split, x = 16) against "
                       "oracle/fen_oracle_c.c stepping the whole grid on rank 0; bound 1e-12" % (nx, ny, nz)}
    dist.barrier()
    G.destroy()
    return out


def run_ns_case(args, cx, case, steps, headline):
    """One navier_stokes_solver benchmark: case "tgv" (BASELINE configs[1], weak scaling) or "channel" (configs[2],
    strong scaling).  Returns the JSON line as a dict (meaningful on rank 0)."""
    import torch
    import torch.distributed as dist
    import fen_b200 as fb
    rank, world, local_rank = cx["rank"], cx["world"], cx["local_rank"]
    n = args.size
    channel = case == "channel"
    if channel:
        # BASELINE configs[2]: turbulent-channel shape 1024 x 1024 x 512 (FEN orientation: walls in z, FFT in x/y,
        # tridiagonal in z), the WHOLE grid split over the ranks (strong scaling); --size scales it down
        nx, ny, nz = 2 * n, 2 * n, n
        Lc = 2.0
        G = fb.grid().setup(nx, ny, nz, Lc, Lc, Lc / 2, pcol=world, rank=rank, device=local_rank,
                            bc=["Periodic"] * 4 + ["Wall", "Wall"])
    else:
        nx, ny, nz = weak_dims(n, world)
        if args.grid:        # tuning aid: an explicit global grid (e.g. the 1024x1024x128 slab one GPU owns at N = 8)
            nx, ny, nz = [int(v) for v in args.grid.split(",")]
        L = 2 * PI
        if args.grid:
            n = min(nx, ny, nz)      # every box side stays a whole number of Taylor-Green wavelengths
        G = fb.grid().setup(nx, ny, nz, L * nx / n, L * ny / n, L * nz / n, pcol=world, rank=rank, device=local_rank)
    if world > 1:
        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        G.connect(all_gather)
    ns = fb.Solver(G, 1.0, 0.01 if not channel else 1.0e-3)
    if channel:
        ns.g = [1.0, 0.0, 0.0]
    ns.init_solver()
    ns.CFL = 0.25
    dt = ns.set_timestep(1.0 if not channel else 1.5)
    if channel:
        keep, (u, v, w, p) = init_channel_slab(G.nloc, (nx, ny, nz), G.delta, G.lo[2] - 1, pinned=True)
    else:
        keep, (u, v, w, p) = init_tgv3d_slab(G.nloc, G.delta, G.lo[2] - 1, pinned=True)
    ns.v.x.f, ns.v.y.f, ns.v.z.f, ns.p.f = u, v, w, p
    ns.v.push(); ns.p.push()
    ns.v.update_ghost_nodes(); ns.p.update_ghost_nodes()
    G.synchronize()
    stream = torch.cuda.ExternalStream(G.lib.fen_gpu_stream(G.ctx))
    ncell = nx * ny * nz
    barrier, timed = _make_timers(G, torch, dist, world, stream)
    step_no = [0]

    def dev_step(_):
        step_no[0] += 1
        ns.navier_stokes_solver(step_no[0], dt)

    for s in range(max(args.warmup, 3)):
        dev_step(s)
    l0 = ns.launch_count()
    with ClockSampler(local_rank) as cs:
        ms = timed(dev_step, steps)
    launches = ns.launch_count() - l0
    clocks = cs.summary()
    maxdiv, maxcfl = ns.status()
    value = ncell * steps / (ms * 1e-3) / 1e6

    # ---- per-kernel timing with CUDA events on the launching stream (same steps, profiled) ----
    barrier()
    nprof = min(steps, 5)
    ns.profile(True)
    for s in range(nprof):
        dev_step(s)
    prof = ns.profile_read()
    ns.profile(False)
    peak, peak_src = load_peaks()
    ncell_loc = G.nloc[0] * G.nloc[1] * G.nloc[2]
    kernels = []
    for name, (tms, cnt) in prof.items():
        if cnt == 0:
            continue
        per = tms / cnt
        bpc = KERNEL_BYTES_PER_CELL.get(name)
        gbs = (bpc * ncell_loc / (per * 1e-3) / 1e9) if bpc else None
        # traffic_frac: bytes the kernel really moved (ncu dram read + write of the committed full-set capture of this
        # command, profiles/ncu_traffic.json) / time / peak -- never above 1, unlike `frac` of a fused kernel, which is
        # credited the algorithmic bytes of everything it replaces
        tb = load_traffic(name, n) if (world == 1 and not channel and not args.grid) else None
        kernels.append({"kernel": name, "launches_per_step": cnt / nprof, "ms_per_launch": per,
                        "ms_per_step": tms / nprof, "alg_bytes_per_cell": bpc,
                        "achieved_GBs": gbs, "frac": (gbs / peak) if gbs else None,
                        "traffic_bytes": tb, "traffic_GBs": (tb / (per * 1e-3) / 1e9) if tb else None,
                        "traffic_frac": (tb / (per * 1e-3) / 1e9 / peak) if tb else None})
    kernels.sort(key=lambda q: -q["ms_per_step"])
    poisson_ms = sum(q["ms_per_step"] for q in kernels
                     if q["kernel"].startswith(("fft_", "thomas_", "mean_line", "a2a_")))
    dom = next((q for q in kernels if q["achieved_GBs"]), None)
    roofline = None
    if dom:
        step_gbs = STEP_BYTES_PER_CELL * ncell_loc / (ms / steps * 1e-3) / 1e9
        tsum = sum(q["traffic_bytes"] for q in kernels if q["traffic_bytes"]) if all(
            q["traffic_bytes"] for q in kernels if q["alg_bytes_per_cell"]) else None
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBs"], "peak": peak,
                    "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic_bytes"],
                    "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                    "peak_source": peak_src,
                    "alg_bytes_per_launch": dom["alg_bytes_per_cell"] * ncell_loc,
                    "step_achieved": step_gbs, "step_frac": step_gbs / peak,
                    "step_traffic_bytes": tsum,
                    "step_traffic_frac": (tsum / (ms / steps * 1e-3) / 1e9 / peak) if tsum else None}

    nvlink = nvlink_figures(args, {q["kernel"]: q["ms_per_step"] for q in kernels}, nx, ny, nz, world,
                            nccl=headline)

    # ---- end to end: host (pinned) arrays in, host arrays out, every step -----------------------
    e2e = None
    if headline and not args.no_e2e:
        fields = [ns.v.x, ns.v.y, ns.v.z, ns.p]
        nbytes = sum(f.f.nbytes for f in fields)

        def e2e_step(_):
            for f in fields:
                f.push()               # waits on the device for the previous step's download of the same array
            step_no[0] += 1
            ns.navier_stokes_solver(step_no[0], dt)
            ns.status()                # the step's scalar result (maxdiv, maxCFL): waits for the uploads and the step
            for f in fields:
                f.pull_async()         # own copy stream: PCIe is full duplex, the next uploads overlap these

        ns.v.pull(); ns.p.pull()
        e2e_step(0)
        ms_e = timed(e2e_step, args.e2e_steps, drain=True)
        e2e = {"value": ncell * args.e2e_steps / (ms_e * 1e-3) / 1e6, "unit": "Mcell-updates/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 16, "ms_per_step": ms_e / args.e2e_steps,
               "what": "push u,v,w,p from pinned host arrays + navier_stokes_solver + status + pull u,v,w,p, every "
                       "step; the pulls run on their own copy stream in pieces, and the next step's upload of an "
                       "array follows its download piece by piece (full-duplex PCIe); the region ends when the last "
                       "download is on the host (dv_o stays on the device: no driver touches it)"}
    barrier()
    G.destroy()
    del keep

    cpu = None
    if headline:
        if args.affinity_before:
            os.sched_setaffinity(0, args.affinity_before)       # the CPU baseline sees every core
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            if channel:
                _, sec, _ = cpu_port_channel(64, 1, 1)
                ksteps = int(max(3, min(40, 15.0 / max(sec, 1e-3))))
                val, sec, cores = cpu_port_channel(64, ksteps, 1)
                cpu = {"value": val, "unit": "Mcell-updates/s", "cores": cores, "kind": "port",
                       "sample": "%d steps of the same channel case at 128x128x64 (1 warm-up), plain-C/OpenMP "
                                 "restatement (oracle/fen_oracle_c.c)" % ksteps}
            else:
                cpu = cpu_baseline(n=args.cpu_size)

    return {
        "metric": "NS timestep Mcell-updates/s", "value": value, "unit": "Mcell-updates/s",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "strong" if channel else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": ("channel %dx%dx%d fp64, walls in z, ppn Poisson: FFT x/y + Thomas z "
                                "(BASELINE configs[2])" % (nx, ny, nz)) if channel else
                               ("3D periodic Taylor-Green vortex %d^3 per GPU fp64, ppp FFT Poisson "
                                "(BASELINE configs[1])" % n), "grid": [nx, ny, nz],
                   "decomposition": "z-slabs x%d" % world,
                   "nu": 1.0e-3 if channel else 0.01, "CFL": 0.25, "dt": dt,
                   "l2": "working set (>= 12 GB) exceeds the 126 MB L2; no flush needed",
                   "cpu_affinity": args.affinity},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "kernels": kernels, "nvlink": nvlink,
        "poisson_solve_ms": poisson_ms,
        "poisson_solve_ms_what": "sum of the solver's kernels in the profiled steps" + (
            " (several ranks: the solve runs in one piece on one stream there; the overlapped solve time is "
            "extra.poisson_only.value)" if world > 1 else ""),
        "check": {"maxdiv": maxdiv, "maxCFL": maxcfl},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cells per direction per GPU (config 2: 512)")
    ap.add_argument("--cpu-size", type=int, default=512,
                    help="CPU arm / cpu_baseline grid (512 = the GPU arm's own configuration)")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline line only: skip the Poisson-only / channel / two-phase / parity runs that the default "
                         "line carries under `extra`")
    ap.add_argument("--no-nccl-baseline", action="store_true",
                    help="N > 1: skip the NCCL all_to_all_single timing reported beside the fused transposes")
    ap.add_argument("--case", default="tgv", choices=["tgv", "channel", "wave2d"],
                    help="tgv: BASELINE configs[1] (headline, weak scaling); channel: configs[2], 2n x 2n x n walls in z")
    ap.add_argument("--grid", default="", help="explicit global grid nx,ny,nz for --case tgv (tuning aid)")
    ap.add_argument("--mode", default="ns", choices=["ns", "poisson"],
                    help="ns: full navier_stokes_solver step (headline); poisson: solve_poisson only (config 4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    cx = _ctx()
    rank, world, local_rank = cx["rank"], cx["world"], cx["local_rank"]
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    args.affinity_before, args.affinity = bind_to_gpu_cpus(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE %d (launch with torch.distributed.run)" % (args.gpus, world))

    if args.case == "wave2d":
        if world != 1:
            raise SystemExit("bench.py: --case wave2d is a one-GPU case (2-D grids are not decomposed)")
        print(json.dumps(run_wave2d(args, local_rank, headline=True)))
        return
    if args.mode == "poisson":
        line = run_poisson_case(args, cx, args.steps)
    else:
        line = run_ns_case(args, cx, args.case, args.steps, headline=True)
        if args.case == "tgv" and not args.grid and not args.no_extras:
            # The other BASELINE configurations, measured by the same command so that the driver's records hold them:
            # never allowed to cost the headline line
            extra = {}
            ksteps = min(args.steps, 10)

            def attempt(key, fn):
                try:
                    extra[key] = fn()
                except BaseException as exc:                      # noqa: BLE001
                    extra[key] = {"unavailable": repr(exc)[:300]}
                    if world > 1:
                        raise                                   # a rank that drops out would hang the others
            attempt("poisson_only", lambda: run_poisson_case(args, cx, ksteps))          # configs[3]
            if world > 1:
                attempt("channel", lambda: run_ns_case(args, cx, "channel", ksteps, headline=False))   # configs[2]
                attempt("parity", lambda: parity_vs_oracle(args, cx, line["config"]["grid"][1],
                                                           line["config"]["grid"][2]))
                if rank == 0 and isinstance(extra.get("parity"), dict) and "rel_l2_vs_oracle" in extra["parity"]:
                    line["check"]["rel_l2_vs_oracle"] = extra["parity"]["rel_l2_vs_oracle"]
            else:
                attempt("wave2d", lambda: run_wave2d(args, local_rank, headline=False))   # configs[4]'s 2-D analogue
                attempt("cufft_crosscheck", lambda: cufft_crosscheck(args.size))
            line["extra"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
