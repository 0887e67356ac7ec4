#!/usr/bin/env python
"""bench.py -- CoLoRe density-field -> sources hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU code on the host cores

One "step" = one pass of the hot path over one synthetic realisation: Gaussian mode fill -> two 3-D
c2r FFTs (+ scaling, sigma^2) -> lognormal transform -> density normalisation -> Poisson sources,
placement, RSD, base pixel, spherical properties. metric = Mcells/s = n_grid^3 / step time.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcells/s end-to-end field->sources"
UNIT = "Mcells/s"
MEAN_SRC_PER_CELL = 0.03          # <sources/cell> ~ 0.03


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_config(n_grid: int):
    from colore_b200.inputs import RunConfig
    cfg = RunConfig(n_grid=n_grid, dens_type=0, seed=1003, n_srcs=1)
    # N(z) amplitude such that the catalogue holds ~0.03 sources per cell (A=3000 -> 948496 objects)
    cfg.nz_amplitude = 3000.0 * MEAN_SRC_PER_CELL * n_grid ** 3 / 948496.0
    return cfg


def build_tables(cfg):
    import colore_b200 as cb
    from colore_b200.inputs import write_inputs
    d = tempfile.mkdtemp(prefix="clr_bench_in_")
    try:
        paths = write_inputs(d, cfg)
        k, pk = np.loadtxt(paths["pk"], unpack=True)
        z, nz = np.loadtxt(paths["nz0"], unpack=True)
        _, bz = np.loadtxt(paths["bz0"], unpack=True)
    finally:
        shutil.rmtree(d)
    return cb.cosmo.cosmo_set(cfg, k, pk, [(z, nz)], [(z, bz)])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index, self.mark_at = None, [], index, 0

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.mark_at = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.mark_at:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref/CoLoRe_ref) on the host cores
def run_reference_sample(n_grid: int, threads: int):
    """Run the reference binary once on a bounded sample; return (Mcells/s, stage dict)."""
    from colore_b200.inputs import write_inputs, write_param_file
    exe = os.path.join(ROOT, "oracle", "_ref", "CoLoRe_ref")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    cfg = make_config(n_grid)
    tmp = tempfile.mkdtemp(prefix="clr_ref_")
    try:
        paths = write_inputs(os.path.join(tmp, "in"), cfg)
        write_param_file(os.path.join(tmp, "param.cfg"), cfg, paths, os.path.join(tmp, "out"))
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        out = subprocess.run([exe, os.path.join(tmp, "param.cfg")], env=env, capture_output=True, text=True,
                             cwd=tmp, timeout=3600).stdout
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    # the reference's own timer() lines (common.c:114-168), in the order main.c runs the stages
    stages, label = {}, None
    keys = [("Creating Fourier-space", "fill"), ("Transforming density", "fft"), ("Normalizing density", "scale"),
            ("Creating physical matter density", "density"), ("Computing normalization", "normalization"),
            ("Getting point sources", "sources"), ("Re-distributing sources", "distribute"),
            ("Writing source catalogs", "write")]
    for ln in out.splitlines():
        for pat, name in keys:
            if pat in ln:
                label = name
        m = re.search(r"Relative time ellapsed\s+([0-9.]+) ms", ln)
        if m and label:
            stages[label] = stages.get(label, 0.0) + float(m.group(1))
            label = None if label != "density" else None
    path_ms = sum(stages.get(k, 0.0) for k in ("fill", "fft", "scale", "density", "normalization", "sources", "distribute"))
    if path_ms <= 0:
        raise RuntimeError("could not parse the reference's timer output:\n" + out[-2000:])
    return n_grid ** 3 / (path_ms * 1e-3) / 1e6, stages


def scipy_fft_ms(n_grid: int, threads: int):
    """Two single-precision c2r transforms of the sample size with scipy's pocketfft on all threads: a sanity figure
    next to the reference's FFT stage, which here runs the oracle shim FFT instead of FFTW."""
    try:
        import scipy.fft as sf
        a = (np.random.default_rng(0).standard_normal((n_grid, n_grid, n_grid // 2 + 1)) + 0j).astype(np.complex64)
        sf.irfftn(a, s=(n_grid,) * 3, workers=threads)
        t0 = time.perf_counter()
        for _ in range(2):
            sf.irfftn(a, s=(n_grid,) * 3, workers=threads)
        return (time.perf_counter() - t0) * 1e3
    except Exception:  # noqa: BLE001
        return None


def workload_config(n: int):
    """`config` of both arms (identical keys and values, so that the driver can tell the two lines describe one workload)."""
    return {"workload": f"n_grid={n} lognormal + 1 galaxy population + RSD (field->sources)", "n_grid": n,
            "mean_sources_per_cell": MEAN_SRC_PER_CELL,
            "l2_policy": f"inputs ({8.0 * n * n * (n // 2 + 1) * 2 / 1e9:.1f} GB of grids) exceed the 126 MB L2",
            "seed_per_step": "varies"}


def reference_arm(args):
    """The unmodified reference on the host cores, SAME n_grid as the GPU arm (a step = one whole run of the binary;
    ~35 s at n_grid=1024 on 16 cores). Steps are capped by a wall-clock budget so that the arm always ends."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_s = args.ref_n_grid or args.n_grid
    # ~35 s per run at n_grid=1024 on 16 cores: a 7-minute budget (warm-up included) keeps the arm "a few minutes" long
    # whatever --steps asks for; the line says how many runs were averaged
    budget_s = float(os.environ.get("CLR_REF_BUDGET_S", "420"))
    try:
        t_start = time.time()
        vals, stages, done = [], {}, 0
        for w in range(1 if args.warmup > 0 else 0):
            run_reference_sample(n_s, threads)
        t0 = time.time()
        for _ in range(args.steps):
            v, stages = run_reference_sample(n_s, threads)
            vals.append(v)
            done += 1
            per = (time.time() - t0) / done
            if time.time() - t_start + per > budget_s:
                break
        ms = (time.time() - t0) * 1e3 / done
        val = float(np.mean(vals))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0][:200]}))
        return
    sample = (f"unmodified reference (oracle/_ref/CoLoRe_ref, gcc -O3 -fopenmp, FFTW replaced by the oracle shim FFT) "
              f"at n_grid={n_s}, same cosmology/tables recipe, {threads} OpenMP threads; stage timers of common.c:114-168 "
              f"summed from mode fill to source redistribution; {done} whole runs of the binary"
              + ("" if done == args.steps else f" (of {args.steps}: {budget_s:.0f} s wall-clock budget)"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.n_grid), "steps_run": done,
        "sample_n_grid": n_s,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                         "stages_ms": stages, "scipy_irfftn_2x_ms": scipy_fft_ms(n_s, threads)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def run_step(cb, par, seed, tabs):
    par.seed = seed
    cb.create_cartesian_fields(par)
    cb.compute_physical_density_field(par)
    cb.compute_density_normalization(par)
    return cb.srcs_set_cartesian(par)[0]


class _DevArray:
    """Raw device pointer -> torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def grid_view(torch, par, which, local):
    pitch = par.grid_pitch()
    t = torch.as_tensor(_DevArray(par.grid_device_ptr(which), par.nz_here * par.n_grid * pitch), device=f"cuda:{local}")
    return t.view(par.nz_here, par.n_grid, pitch)


PARITY_SEED = 4242
# catalogue length of the PARITY_SEED realisation of the bench workload at n_grid=1024 on ONE GPU (measured, round 2):
# the Poisson stream is keyed by the global cell index, so every slab decomposition must reproduce it
PARITY_NSRC_1024 = 31129449


def parity_checks(cb, torch, par, tabs, n, local, allsum, allmax):
    """Correctness bits carried by every bench line (size-independent properties, cheap): distributed r2c(c2r(.))
    round trip of the potential, sum of the per-cell counts == catalogue length, total number of sources of a fixed seed
    (identical for every number of GPUs), moments of the Gaussian field equal on every rank."""
    out = {}
    par.seed = PARITY_SEED
    mean, s2 = cb.create_cartesian_fields(par)
    par.synchronize()
    out["sigma2_gauss"] = s2
    out["sigma2_same_on_all_ranks"] = bool(allmax(s2) == s2 and -allmax(-s2) == s2)
    g = grid_view(torch, par, cb.GRID_NPOT, local)
    ref = g[:, :, :n].clone()
    torch.cuda.synchronize()
    cb.fftw_wrap_r2c(par, cb.GRID_NPOT)
    cb.fftw_wrap_c2r(par, cb.GRID_NPOT)
    par.synchronize()
    err = float((g[:, :, :n] * (1.0 / float(n) ** 3) - ref).abs().max())
    s1 = allsum(float(ref.double().sum()))
    s2r = allsum(float((ref.double() ** 2).sum()))
    sig = (s2r / float(n) ** 3 - (s1 / float(n) ** 3) ** 2) ** 0.5
    out["fft_roundtrip_err_over_sigma"] = allmax(err) / sig
    g[:, :, :n].copy_(ref)
    del ref
    torch.cuda.synchronize()
    par.update_halo()
    cb.compute_physical_density_field(par)
    cb.compute_density_normalization(par)
    nsrc = cb.srcs_set_cartesian(par)[0]
    counts = cb.srcs_get_counts(par, 0)
    out["sum_counts_equals_catalogue"] = bool(allsum(float(int(counts.sum()) == nsrc)) == allsum(1.0))
    tot = int(allsum(nsrc))
    out["sources_total_parity_seed"] = tot
    if n == 1024 and PARITY_NSRC_1024 is not None:
        out["sources_total_equals_1gpu_value"] = bool(tot == PARITY_NSRC_1024)
    out["ok"] = bool(out["fft_roundtrip_err_over_sigma"] < 5e-5 and out["sum_counts_equals_catalogue"]
                     and out["sigma2_same_on_all_ranks"] and out.get("sources_total_equals_1gpu_value", True))
    return out


def ncu_traffic_table():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed `ncu --set full` summaries of THIS
    workload (n_grid=1024, one GPU) under profiles/ (written by tools/ncu_summary.py)."""
    import csv
    import glob
    stage_of = {"fill_z_kernel": "fill_fft_z", "yx_fused_kernel": "fft_yx", "fill_modes_fast_kernel": "fill_modes",
                "lognormal_hist_kernel": "lognormal", "lognormal_fast_kernel": "lognormal", "norm_hist_fast_kernel": "norm_hist",
                "poisson_kernel": "srcs_poisson", "expand_kernel": "srcs_expand", "place_src_kernel": "srcs_place"}
    tab, src = {}, {}
    def order(f):                                  # round, then the _vN tag of the capture: later captures override earlier ones
        b = os.path.basename(f)
        v = re.search(r"_v(\d+)_", b)
        return (int(re.match(r"r(\d+)_", b).group(1)), int(v.group(1)) if v else 0, b)
    files = [f for f in glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*_full.csv"))
             if "_1024_" in os.path.basename(f) and re.match(r"r\d+_", os.path.basename(f))]
    for f in sorted(files, key=order):
        try:
            for row in csv.DictReader(open(f)):
                name = row.get("kernel", "")
                for key, st in stage_of.items():
                    if key in name and "dram__bytes_read.sum" in row:
                        def gb(x):
                            v, u = x.split()[:2]
                            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                        tab[st] = gb(row["dram__bytes_read.sum"]) + gb(row["dram__bytes_write.sum"])
                        src[st] = os.path.basename(f)
                        if key == "norm_hist_fast_kernel" and "1, 1>" in name.replace("1, 1, 2>", "1, 1>"):
                            tab["lognormal"], src["lognormal"] = tab[st], src[st]     # the fused lognormal + histogram walk
        except Exception:  # noqa: BLE001
            continue
    return tab, src


def north_star_run(cb, torch, dist, rank, world, local, allsum, allmax, barrier, steps=3):
    """BASELINE.json north star: 2048^3 lognormal + sources + kappa (nside 1024) on 8 GPUs, with its property checks and
    the HBM fractions of the FFT / lognormal stages. Appended to the --gpus 8 line under "north_star"."""
    n = 2048
    cfg = make_config(n)
    tabs = build_tables(cfg)
    nzl, iz0 = cb.dist.slab_bounds(n, world, rank)
    par = cb.ParamCoLoRe(tabs, n, dens_type=0, seed=cfg.seed, device=local, nz_here=nzl, iz0_here=iz0)
    cb.dist.init_comm(par, rank, world)
    par.set_srcs(0, tabs["srcs_nz_0"], tabs["srcs_bz_0"])
    out = {"workload": "n_grid=2048 lognormal + 1 galaxy population + RSD + kappa/ISW nside 1024", "n_gpus": world,
           "transpose": cb.dist.transpose_mode(par)}
    out["checks"] = parity_checks(cb, torch, par, tabs, n, local, allsum, allmax)
    run_step(cb, par, 7, tabs)
    par.synchronize()
    barrier()
    par.set_profiling(True)
    par.timer_start()
    nsrc = 0
    for s in range(steps):
        nsrc = run_step(cb, par, 100 + s, tabs)
    ms = allmax(par.timer_stop_ms()) / steps
    st = {}
    for nm in STAGES:
        m, nl = par.stage_ms(nm)
        if m or nl:
            st[nm] = {"ms_per_step_max": allmax(m) / steps, "ms_per_step_min": -allmax(-m) / steps}
    par.set_profiling(False)
    peak, _ = measured_hbm_peak()
    cells = float(n) ** 3 / world
    fft_ms = sum(st[k]["ms_per_step_max"] for k in ("fill_modes", "fill_fft_z", "fft_z", "fft_y", "fft_x", "fft_yx") if k in st)
    out.update({"ms_per_step": ms, "Mcells_per_s": float(n) ** 3 / ms / 1e3, "sources_per_step": int(allsum(nsrc)), "stages": st,
                "fill_plus_fft_hbm_gbs_per_gpu": 56.0 * cells / (fft_ms * 1e-3) / 1e9,
                "fill_plus_fft_frac_of_hbm_peak": 56.0 * cells / (fft_ms * 1e-3) / 1e9 / peak})
    if "lognormal" in st:
        lb = 12.0 if "norm_hist" not in st else 8.0       # lognormal alone: 8 B/cell; with the fused histogram: + 4
        out["lognormal_frac_of_hbm_peak"] = lb * cells / (st["lognormal"]["ms_per_step_max"] * 1e-3) / 1e9 / peak
    if out["transpose"] == "p2p-fused":
        zname = "fill_fft_z" if "fill_fft_z" in st else "fft_z"
        sent = 2 * 8.0 * n * n * (n // 2 + 1) / world * (world - 1) / world
        out["nvlink_gbs_per_direction_during_fused_pass"] = sent / (st[zname]["ms_per_step_max"] * 1e-3) / 1e9
    out.update(maps_run(cb, par, tabs, allmax, barrier))
    par.free()
    return out


def maps_run(cb, par, tabs, allmax, barrier, nside=1024):
    """kappa + ISW HEALPix maps (kappa.c:39-175, isw.c:78-147) at ``nside`` on two source planes from the fields ``par``
    holds: kernel time (max over ranks; kappa includes its Hessian precompute pass) and wall time of the API call with
    the host pixel list going in and the all-reduced map coming out. BASELINE config 3 at the bench's n_grid."""
    out = {}
    rf = np.interp([0.2, 0.4], tabs["z"], tabs["r"]).astype(np.float32)
    _, pos = cb.healpix.hp_shell_pixels(nside, 2)
    for name, fn in (("kappa_los", cb.kappa_get_beam_properties), ("isw_los", cb.isw_get_beam_properties)):
        fn(par, pos[:1024], rf)
        barrier()
        par.set_profiling(True)
        t0 = time.perf_counter()
        m = fn(par, pos, rf)
        wall = allmax(time.perf_counter() - t0)
        kms, _ = par.stage_ms(name)
        if name == "kappa_los":
            kms += par.stage_ms("kappa_tidal")[0]           # the Hessian precompute pass belongs to the kappa stage
        par.set_profiling(False)
        out[name] = {"nside": nside, "kernel_ms_max": allmax(kms), "api_wall_ms": wall * 1e3,
                     "finite": bool(np.isfinite(m).all()), "map_rms": float(m.astype(np.float64).std())}
    return out


STAGES = ["fill_modes", "fill_fft_z", "fft_z", "fft_a2a", "fft_y", "fft_x", "fft_yx", "halo", "lognormal", "norm_hist",
          "srcs_bound", "srcs_poisson", "srcs_scan", "srcs_expand", "srcs_place", "srcs_local"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-grid", type=int, default=1024)
    ap.add_argument("--ref-n-grid", type=int, default=0,
                    help="n_grid of the CPU runs: --impl reference defaults to --n-grid (same config); the cpu_baseline leg "
                         "of the GPU arm defaults to 512 (a bounded sample)")
    ap.add_argument("--no-north-star", action="store_true", help="skip the 2048^3 run that --gpus 8 appends")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-maps", action="store_true", help="skip the kappa / ISW maps (nside 1024) appended under \"maps\"")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="clr_set_option before the run (kernel-variant experiments, e.g. --opt fill_w=4)")
    ap.add_argument("--writer", action="store_true",
                    help="also time clr_write_catalog (ASCII + FITS, all host threads) on the last catalogue; off by default "
                         "because it writes ~2.4 GB into the temporary directory")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch

    import colore_b200 as cb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n_grid
    cfg = make_config(n)
    tabs = build_tables(cfg)
    # strong scaling: the n_grid^3 box is cut into `world` z slabs (fourier.c:172-177), one per GPU
    nz_here, iz0_here = cb.dist.slab_bounds(n, world, rank)
    par = cb.ParamCoLoRe(tabs, n, dens_type=0, seed=cfg.seed, device=local, nz_here=nz_here, iz0_here=iz0_here)
    cb.dist.init_comm(par, rank, world)
    nz_tab, bz_tab = tabs["srcs_nz_0"], tabs["srcs_bz_0"]
    par.set_srcs(0, nz_tab, bz_tab)
    for o in args.opt:
        par.set_option(o.split("=")[0], int(o.split("=")[1]))

    # ---- device-resident timing (value) ------------------------------------------------------
    # nvidia-smi takes ~100 ms to deliver its first line: start it before the warm-up and count only the
    # samples taken from the start of the timed regions (device-timed loop + end-to-end loop) on
    sampler = ClockSampler(local)
    sampler.start()
    for w in range(args.warmup):
        run_step(cb, par, 100 + w, tabs)
    par.synchronize()
    barrier()
    sampler.mark()
    par.set_profiling(True)
    l0 = par.launch_count
    par.timer_start()
    nsrc = 0
    for s in range(args.steps):
        # inputs (two 4.3 GB grids at 1024^3) are far larger than the 126 MB L2: no flush needed
        nsrc = run_step(cb, par, 1000 + s, tabs)
    ms_total = par.timer_stop_ms()
    barrier()
    ms_total = allmax(ms_total)                      # device time, max over ranks
    nsrc_total = int(allsum(nsrc))
    launches = par.launch_count - l0
    ms_step = ms_total / args.steps
    value = n ** 3 / (ms_step * 1e-3) / 1e6
    stages = {}
    for nm in STAGES:
        ms, nl = par.stage_ms(nm)
        if nl or ms:
            # this rank's value (rank 0 prints) + min / max over the ranks: edge slabs hold fewer sources than middle ones
            stages[nm] = {"ms_per_step": ms / args.steps, "launches_per_step": nl / args.steps,
                          "ms_per_step_max": allmax(ms) / args.steps, "ms_per_step_min": -allmax(-ms) / args.steps}
    par.set_profiling(False)

    # ---- roofline of the dominant kernel -------------------------------------------------------
    peak, peak_src = measured_hbm_peak()
    nc = n // 2 + 1
    grid_bytes = 8.0 * n * n * nc / world   # this rank's slab of a complex64 half-spectrum / padded real grid
    cells = float(n) ** 3 / world           # cells of this rank's slab
    alg_bytes = {                           # algorithmic bytes per LAUNCH
        "fill_modes": 2 * grid_bytes,                      # two complex grids written (8 B/cell)
        "fft_z": 2 * grid_bytes, "fft_y": 2 * grid_bytes, "fft_x": 2 * grid_bytes,   # 8 B/cell per pass
        # fused kernels, counted as the stages they replace: fill (8 B/cell) + the z pass of both fields (2 x 8 B/cell);
        # y pass + x pass of one field (2 x 8 B/cell)
        "fill_fft_z": 6 * grid_bytes, "fft_yx": 4 * grid_bytes,
        "lognormal": 8.0 * cells, "norm_hist": 4.0 * cells, "srcs_poisson": 8.0 * cells,
        "srcs_expand": 4.0 * cells + 8.0 * nsrc, "srcs_place": 36.0 * nsrc,
    }
    dom = max((k for k in stages if k in alg_bytes and stages[k]["launches_per_step"] > 0),
              key=lambda k: stages[k]["ms_per_step"] / stages[k]["launches_per_step"])
    per_launch_ms = stages[dom]["ms_per_step"] / stages[dom]["launches_per_step"]
    achieved = alg_bytes[dom] / (per_launch_ms * 1e-3) / 1e9
    ncu_traffic, ncu_src = ncu_traffic_table()
    traffic = ncu_traffic.get(dom) if (n == 1024 and world == 1) else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": ncu_src.get(dom) if traffic else None,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "ms_per_launch": per_launch_ms}
    # mode fill + both 3-D c2r transforms: 8 + 2 x 24 B/cell (the fill is fused into the z pass on one GPU)
    fft_ms = sum(stages[k]["ms_per_step"] for k in ("fill_modes", "fill_fft_z", "fft_z", "fft_y", "fft_x", "fft_yx") if k in stages)
    fft_gbs = (8.0 + 2 * 24.0) * cells / (fft_ms * 1e-3) / 1e9 if fft_ms else None        # per GPU
    nvlink = None
    if world > 1 and cb.dist.transpose_mode(par) == "p2p-fused":
        # the transpose is fused into the z pass: its stores go straight to the destination GPU over NVLink,
        # so the bytes below travel DURING that kernel (time = the whole fused pass incl. its two barriers)
        sent = 2 * grid_bytes * (world - 1) / world
        nvlink = {"transpose": "p2p-fused (peer stores from the FFT z pass)", "bytes_sent_per_rank_per_step": sent,
                  "ms_per_step": stages["fft_z"]["ms_per_step"],
                  "achieved_gbs_per_direction": sent / (stages["fft_z"]["ms_per_step"] * 1e-3) / 1e9,
                  "peak_gbs_per_direction": 770.0, "peak_source": "measured peer copy (B200_PROFILING.md)"}
    elif world > 1 and "fft_a2a" in stages and stages["fft_a2a"]["ms_per_step"] > 0:
        # bytes one rank SENDS per step: 2 transforms x slab bytes x (P-1)/P
        sent = 2 * grid_bytes * (world - 1) / world
        nvlink = {"transpose": "nccl all-to-all", "bytes_sent_per_rank_per_step": sent, "ms_per_step": stages["fft_a2a"]["ms_per_step"],
                  "achieved_gbs_per_direction": sent / (stages["fft_a2a"]["ms_per_step"] * 1e-3) / 1e9,
                  "peak_gbs_per_direction": 770.0, "peak_source": "measured peer copy (B200_PROFILING.md)"}

    # ---- end to end through the public API with host buffers -----------------------------------
    # inputs: the population tables from pinned host memory (H2D every step); result: the Src records
    # (common.h:169-179) copied into pinned host memory (D2H every step)
    pin_in = torch.empty(2 * cb._lib.NA, dtype=torch.float64).pin_memory()
    pin_in[:cb._lib.NA] = torch.from_numpy(np.nan_to_num(nz_tab))
    pin_in[cb._lib.NA:] = torch.from_numpy(np.nan_to_num(bz_tab))
    tin = pin_in.numpy()
    cap = int(nsrc * 1.2) + 1024
    # two pinned result buffers: with async_results the read-back of step s overlaps the kernels of step
    # s+1 (the copy engine and the SMs work at the same time); all K results are home when the clock stops
    pin_out = [torch.empty((cap, 9), dtype=torch.float32).pin_memory() for _ in range(2)]
    tout = [p.numpy() for p in pin_out]
    d2h = 0
    # the catalogue copy alone (synchronous), for reference: PCIe sets the floor of an un-overlapped read-back
    t0 = time.perf_counter()
    cb.srcs_get_local_properties(par, 0, out=tout[0][:nsrc])
    d2h_alone_ms = (time.perf_counter() - t0) * 1e3
    par.set_option("async_results", 1)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        par.set_srcs(0, tin[:cb._lib.NA], tin[cb._lib.NA:])
        k = run_step(cb, par, 2000 + s, tabs)
        cb.srcs_get_local_properties(par, 0, out=tout[s & 1][:k])
        d2h += k * 36
    par.synchronize()
    par.set_option("async_results", 0)
    e2e_ms = allmax((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks = sampler.stop()
    e2e = {"value": n ** 3 / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(tin.nbytes * world),
           "d2h_bytes_per_step": int(allsum(d2h / args.steps)), "ms_per_step": e2e_ms,
           "d2h_alone_ms_rank0": d2h_alone_ms,
           "note": "read-back of step s runs on a copy stream under the kernels of step s+1 (two device and two pinned "
                   "host buffers); all K catalogues are home when the clock stops"}

    # ---- CPU baseline (rank 0, bounded sample) ---------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:      # N=1 only (the other ranks would just wait)
        threads = os.cpu_count() or 1
        try:
            n_cpu = args.ref_n_grid or 512
            v, st = run_reference_sample(n_cpu, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"unmodified reference (oracle/_ref/CoLoRe_ref; shim FFT instead of FFTW) at "
                             f"n_grid={n_cpu} (bounded sample; `--impl reference` runs the full n_grid), {threads} OpenMP "
                             f"threads, field->sources stages",
                   "stages_ms": st, "scipy_irfftn_2x_ms": scipy_fft_ms(n_cpu, threads)}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"failed: {e}"[:200]}

    # ---- (optional) the native catalogue writer on the last catalogue --------------------------------
    writer = None
    if args.writer and rank == 0:
        wd = tempfile.mkdtemp(prefix="clr_bench_out_")
        try:
            writer = {"sources": int(par.nsources[0]), "threads": os.cpu_count()}
            for fmt in ("fits", "ascii"):
                fn = os.path.join(wd, "cat." + fmt)
                sec = cb.write_catalog(par, 0, fn, fmt)
                writer[fmt] = {"seconds": sec, "bytes": os.path.getsize(fn), "Msources_per_s": par.nsources[0] / sec / 1e6}
        finally:
            shutil.rmtree(wd, ignore_errors=True)

    # ---- correctness bits of this very configuration (after the timed regions) ------------------
    parity = parity_checks(cb, torch, par, tabs, n, local, allsum, allmax)
    transpose = cb.dist.transpose_mode(par) if world > 1 else "none"
    # ---- BASELINE config 3: kappa + ISW maps at nside 1024 from the fields of this configuration (outside the step) ----
    maps = None
    if not args.no_maps:
        try:
            run_step(cb, par, 1000, tabs)
            maps = maps_run(cb, par, tabs, allmax, barrier)
        except Exception as e:  # noqa: BLE001
            maps = {"failed": str(e)[:300]}
    par.free()
    par = None
    north = None
    if world == 8 and not args.no_north_star:
        try:
            north = north_star_run(cb, torch, dist, rank, world, local, allsum, allmax, barrier)
        except Exception as e:  # noqa: BLE001
            north = {"failed": str(e)[:300]}

    if rank == 0:
        cfgd = workload_config(n)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfgd, "sources_per_step": nsrc_total,
            "parallelism": f"{world} z-slab(s), one process per GPU, FFT slab transpose: {transpose}",
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "fft_hbm_gbs": fft_gbs, "nvlink": nvlink, "stages": stages, "parity": parity,
            "maps": maps, "north_star": north, "writer": writer, "cpu_baseline": cpu,
        }))
    if par is not None:
        par.free()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
