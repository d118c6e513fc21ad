"""CPU-only checks of the product's host side: the C-ABI library builds, loads, exports every symbol
include/fen_gpu.h declares, fails loudly without a device, and the FFT index logic is right."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fen_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from fen_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fen_gpu.h")).read()
    declared = set(re.findall(r"\b(fen_gpu_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fen_gpu_[a-z_0-9]+)", out))
    assert declared <= exported


def test_library_is_sm100a_and_has_no_torch_dependency(lib):
    from fen_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "libcufft" not in ldd


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import fen_b200 as fb
    with pytest.raises(fb.FenError) as e:
        fb.grid().setup(16, 16, 16, 1.0, 1.0, 1.0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fen_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().replace("oracle's", ""), f


def test_fft_core_index_logic_on_cpu():
    exe = os.path.join(ROOT, "build", "test_fft_core")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "cpu", "test_fft_core.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout


def test_sass_carries_the_copy_engine_instructions():
    """The built library's SASS (cuobjdump, no GPU needed): the TMA tensor loads of the stencil kernels (UTMALDG), the bulk
    shared->global copies the slab transposes ship their tiles with (UBLKCP) and the mbarrier operations (SYNCS) are
    really there -- the evidence profiles/r02_sass_summary.txt records, kept from regressing."""
    import shutil
    from fen_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    r = subprocess.run("cuobjdump -sass %s | python %s" % (_lib.LIB_PATH, os.path.join(ROOT, "scripts", "sass_count.py")),
                       shell=True, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = {}
    hdr = None
    for line in r.stdout.splitlines():
        cols = line.split()
        if line.startswith("kernel"):
            hdr = cols[1:]
            continue
        n = len(hdr)
        rows[" ".join(cols[:-n])] = dict(zip(hdr, (int(x) for x in cols[-n:])))
    pick = lambda frag: [v for k, v in rows.items() if frag in k]
    assert pick("k_pred_tma") and all(v["UTMALDG"] >= 3 and v["SYNCS"] >= 1 for v in pick("k_pred_tma"))
    assert pick("k_corr_tma") and all(v["UTMALDG"] >= 4 for v in pick("k_corr_tma"))
    for frag in ("k_fft_lines_bs<1024", "k_fft_solve_bs<1024", "k_fft_lines_bs<512", "k_fft_solve_bs<512", "k_bulk_rows"):
        assert pick(frag) and all(v["UBLKCP"] >= 1 for v in pick(frag)), frag
    assert all(v["LDGSTS"] > 0 for v in pick("k_thomas_lp"))


def test_any_length_kernels_run_on_cpu():
    """fen_b200/csrc/fft_any.cuh: the any-length Poisson kernels (non-power-of-two grids) are sequences of
    __host__ __device__ phases; tests/cpu/test_fft_any.cu runs them with blocks and threads as loops against direct
    O(n^2) DFT / DCT-II / DCT-III sums -- strided lines (forward, backward, fused solve with the singular mode), r2c /
    c2r rows with the periodic ghosts, cosine transforms, lengths 1 ... 6144 with factors 2, 3, 5, 7, 11, 13, 29, 31, 37, 53, 61."""
    exe = os.path.join(ROOT, "build", "test_fft_any")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(ROOT, "tests", "cpu", "test_fft_any.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK") and "FAIL" not in r.stdout, r.stdout[-2000:]


def _build_c_driver(name="tgv_driver"):
    exe = os.path.join(ROOT, "build", name)
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    from fen_b200 import _lib
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", name + ".c"), "-L", libdir, "-lfen_gpu",
                    "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    return exe


def _build_cpp_driver(name="tgv_driver"):
    exe = os.path.join(ROOT, "build", name + "_cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    from fen_b200 import _lib
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", name + ".cpp"), "-L", libdir, "-lfen_gpu",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_cpp_mirror_compiles_and_a_cpp_driver_links(lib):
    """include/fen_gpu.hpp -- the header-only C++ mirror of the reference's solver API (grid%setup, scalar / vector,
    init_solver, set_timestep, advance_solution, print_solver_status ...) -- compiles with -Wall -Wextra -pedantic
    -Werror, the Taylor-Green driver written against it links, and without a device it stops at grid%setup with the 'no
    CPU fallback' message (the run itself is tests/test_gpu_parity.py::test_cpp_driver_runs)."""
    import torch
    exe = _build_cpp_driver()
    if torch.cuda.is_available():
        pytest.skip("a device is present: the run itself is tests/test_gpu_parity.py::test_cpp_driver_runs")
    r = subprocess.run([exe, "16", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_two_phase_c_driver_links(lib):
    """examples/shear_drop_driver.c: the reference's shear-drop driver (-DMF build) through the two-phase entry points
    of the C ABI -- module parameters, init_solver with a distance callback, moving-wall values -- compiles as C99 with
    -pedantic -Werror and links; without a device it stops at fen_gpu_create (the run is a GPU test)."""
    import torch
    exe = _build_c_driver("shear_drop_driver")
    if torch.cuda.is_available():
        pytest.skip("a device is present: the run itself is tests/test_gpu_zz_rising_bubble.py")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_c_abi_header_is_plain_c_and_a_c_driver_links(lib):
    """include/fen_gpu.h must be consumable from C (the Fortran shim binds the same symbols): a C99 driver with the
    reference's Taylor-Green call sequence compiles with -pedantic -Werror and links; without a device it stops at
    fen_gpu_create with the 'no CPU fallback' message."""
    import torch
    exe = _build_c_driver()
    if torch.cuda.is_available():
        pytest.skip("a device is present: the run itself is tests/test_gpu_parity.py::test_c_driver_runs")
    r = subprocess.run([exe, "16", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_ctypes_signatures_match_the_header_prototypes():
    """fen_b200/_lib.py: every argtypes list has the length of the C prototype, by-value ints / doubles are c_int /
    c_double in the same positions, and pointer parameters are pointer-like ctypes."""
    import ctypes as C
    from fen_b200 import _lib
    from tests.test_fortran_shim import c_prototypes
    protos = c_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (res, argtypes) in _lib.SIGNATURES.items():
        params = protos[name]
        assert len(params) == len(argtypes), (name, params, argtypes)
        for (is_ptr, ctext), at in zip(params, argtypes):
            if is_ptr:
                assert at in (C.c_void_p, C.c_char_p) or hasattr(at, "contents") or hasattr(at, "_flags_"), (name, ctext, at)
            elif re.match(r"(const\s+)?double\b", ctext):
                assert at is C.c_double, (name, ctext, at)
            else:
                assert at is C.c_int, (name, ctext, at)


def test_grid_print_json_has_the_reference_format(tmp_path):
    """grid%print_json (grid.f90:233-264): the file every postpro.py of the reference opens first (e.g.
    test/large_test/lid3D/postpro.py:20-35).  Formats I7 / E16.8, line for line; json.load must read it back."""
    import json
    import math
    from fen_b200 import api
    text = api.grid_json(64, 128, 1, (0.0, -0.5, 0.0), 1.0, 2.0, 1.0 / 64)
    lines = text.splitlines()
    assert lines[0] == "{" and lines[1] == '    "Grid": {' and lines[-2] == "      }" and lines[-1] == "}"
    assert lines[2] == '        "Nx":       64,' and lines[3] == '        "Ny":      128,'
    assert lines[5] == '        "origin": [   0.00000000E+00, -0.50000000E+00,  0.00000000E+00],'
    assert lines[6] == '        "Lx":    0.10000000E+01,' and lines[8] == '        "Lz":    0.15625000E-01'
    g = json.loads(text)["Grid"]
    assert (g["Nx"], g["Ny"], g["Nz"], g["Lx"], g["Ly"]) == (64, 128, 1, 1.0, 2.0) and g["origin"][1] == -0.5
    assert api._fortran_e(2.0 * math.pi) == "  0.62831853E+01" and api._fortran_e(0.999999999) == "  0.10000000E+01"
    # the method writes <name>.json on rank 0 only (no device needed: a bare instance with the grid attributes)
    G = api.grid()
    G.Nx, G.Ny, G.Nz, G.origin, G.Lx, G.Ly, G.Lz, G.rank = 16, 16, 16, (0.0, 0.0, 0.0), 1.0, 1.0, 1.0, 0
    path = G.print_json(str(tmp_path))
    assert os.path.basename(path) == "grid.json" and json.load(open(path))["Grid"]["Nz"] == 16
    G.rank = 1
    assert G.print_json(str(tmp_path)) is None


def test_closest_grid_node_matches_the_reference_rule():
    """grid%closest_grid_node (grid.f90:204-229, exercised by test/small_test/grid/main.f90): minloc over the points of
    the requested location; 1-based; no device needed."""
    import numpy as np
    from fen_b200 import api
    G = api.grid()
    G.Nx, G.Ny, G.Nz, G.ndim, G.delta = 8, 4, 2, 3, 0.25
    G.x = -1.0 + (np.arange(0, G.Nx + 2) - 0.5) * G.delta
    G.y = 0.0 + (np.arange(0, G.Ny + 2) - 0.5) * G.delta
    G.z = 0.0 + (np.arange(0, G.Nz + 2) - 0.5) * G.delta
    assert G.closest_grid_node([-0.9, 0.6, 0.3], 0) == [1, 3, 2]          # cell centres at -0.875, 0.625, 0.375
    assert G.closest_grid_node([-0.76, 0.6, 0.3], 1) == [1, 3, 2]         # x faces at -0.75, -0.5, ...
    assert G.closest_grid_node([-0.76, 0.74, 0.3], 2) == [1, 3, 2]        # y faces at 0.25, 0.5, 0.75, 1.0
    assert G.closest_grid_node([100.0, -100.0, 0.0], 4) == [8, 1, 1]
    G.ndim = 2
    assert G.closest_grid_node([-0.9, 0.6], 0) == [1, 3, 1]


@pytest.mark.parametrize("case", ["tgv", "channel", "wave2d"])
def test_bench_reference_arm_prints_the_contract_line(case):
    """bench.py --impl reference (the CPU restatement timed on the host cores, no GPU needed): one JSON line with the
    keys of the contract -- metric, value, unit, n_gpus, steps, warmup, ms_per_step, higher_is_better, scaling,
    vs_baseline, dtype, data, config.workload, impl, cpu_baseline{value, unit, cores, kind, sample}, e2e{value, unit,
    h2d_bytes_per_step = d2h_bytes_per_step = 0} -- for each bench case, on a small sample."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--case", case, "--steps",
                        "1", "--warmup", "1", "--cpu-size", "32"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["value"] > 0 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_bench_host_logic():
    """bench.py's pure host logic: the weak-scaling shapes are BASELINE configs[3]'s (512^3 per GPU: x, then y, then z
    doubles), every timed kernel family finds its DRAM traffic in the committed ncu capture (so that no
    kernels[].traffic_frac is left empty or above 1 for lack of a match), and the transposes' figures come out of the
    kernel table as documented."""
    import argparse
    import json
    sys.path.insert(0, ROOT)
    import bench
    assert [bench.weak_dims(512, w) for w in (1, 2, 4, 8)] == [[512, 512, 512], [1024, 512, 512], [1024, 1024, 512],
                                                              [1024, 1024, 1024]]
    with pytest.raises(SystemExit):
        bench.weak_dims(512, 3)
    traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert traffic["size"] == 512
    for fam, lo, hi in (("pred", 104, 112), ("corr_check", 72, 80), ("fft_x_r2c_div", 32, 34), ("fft_solve", 16, 17),
                        ("fft_lines_fwd", 16, 17), ("fft_lines_inv", 16, 17), ("fft_x_c2r", 16, 17)):
        b = bench.load_traffic(fam, 512)
        assert b is not None, fam
        assert lo <= b / 512 ** 3 <= hi, (fam, b / 512 ** 3)            # bytes per cell the kernel really moved
        assert b <= bench.KERNEL_BYTES_PER_CELL[fam] * 512 ** 3 * 1.08  # never far above the algorithmic bytes
    assert bench.load_traffic("pred", 256) is None                     # another size: no figure rather than a wrong one
    args = argparse.Namespace(no_nccl_baseline=True)
    tk = {"fft_lines_fwd_a2a": 1.0, "a2a_fwd_sync": 0.25, "fft_solve_a2a": 1.5, "a2a_bwd_sync": 0.5}
    nv = bench.nvlink_figures(args, tk, 1024, 1024, 1024, 8, nccl=False)
    sent = 513 * 1024 * 128 * 16 * 7 // 8
    assert nv["a2a_bytes_sent_per_gpu"] == sent
    assert abs(nv["fwd_bus_GBs"] - sent / 1.25e-3 / 1e9) < 1e-9 and abs(nv["bwd_bus_GBs"] - sent / 2.0e-3 / 1e9) < 1e-9
    assert bench.nvlink_figures(args, tk, 512, 512, 512, 1) is None
