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


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_s = args.ref_n_grid
    cfg_name = f"n_grid={args.n_grid} lognormal + 1 galaxy population + RSD (field->sources)"
    try:
        for _ in range(args.warmup if args.warmup < 2 else 1):
            run_reference_sample(n_s, threads)
        vals = []
        t0 = time.time()
        for _ in range(args.steps):
            v, stages = run_reference_sample(n_s, threads)
            vals.append(v)
        ms = (time.time() - t0) * 1e3 / args.steps
        val = float(np.mean(vals))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0][:200]}))
        return
    sample = (f"unmodified reference (oracle/_ref/CoLoRe_ref, gcc -O3 -fopenmp, FFTW replaced by the oracle shim FFT) "
              f"at n_grid={n_s}, same cosmology/tables recipe, {threads} OpenMP threads; stage timers of common.c:114-168 "
              f"summed from mode fill to source redistribution")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": cfg_name, "sample_n_grid": n_s},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-grid", type=int, default=1024)
    ap.add_argument("--ref-n-grid", type=int, default=512, help="bounded sample size of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    stage_names = ["fill_modes", "fill_fft_z", "fft_z", "fft_a2a", "fft_y", "fft_x", "fft_yx", "halo", "lognormal", "norm_hist", "srcs_poisson",
                   "srcs_scan", "srcs_expand", "srcs_place", "srcs_local"]
    stages = {}
    for nm in stage_names:
        ms, nl = par.stage_ms(nm)
        if nl or ms:
            stages[nm] = {"ms_per_step": ms / args.steps, "launches_per_step": nl / args.steps}
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
              key=lambda k: stages[k]["ms_per_step"])
    per_launch_ms = stages[dom]["ms_per_step"] / stages[dom]["launches_per_step"]
    achieved = alg_bytes[dom] / (per_launch_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of THIS
    # workload (n_grid=1024, one GPU): profiles/r1_ncu_top_kernels_v8_full.csv (Poisson: ..._v4_full.csv)
    ncu_traffic = {"fill_modes": 9.33e9, "fft_z": 8.55e9, "fft_y": 8.55e9, "fft_x": 8.56e9, "lognormal": 8.58e9,
                   "norm_hist": 4.33e9, "srcs_poisson": 9.32e9, "srcs_expand": 3.72e9, "srcs_place": 5.81e9}
    traffic = ncu_traffic.get(dom) if (n == 1024 and world == 1) else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
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
            v, st = run_reference_sample(args.ref_n_grid, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"unmodified reference (oracle/_ref/CoLoRe_ref; shim FFT instead of FFTW) at "
                             f"n_grid={args.ref_n_grid}, {threads} OpenMP threads, field->sources stages",
                   "stages_ms": st, "scipy_irfftn_2x_ms": scipy_fft_ms(args.ref_n_grid, threads)}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"failed: {e}"[:200]}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"n_grid={n} lognormal + 1 galaxy population + RSD (field->sources)",
                       "n_grid": n, "sources_per_step": nsrc_total,
                       "parallelism": f"{world} z-slab(s), one process per GPU, FFT slab transpose: "
                                      + (cb.dist.transpose_mode(par) if world > 1 else "none"),
                       "l2_policy": "inputs (8.6 GB of grids at 1024^3) exceed the 126 MB L2", "seed_per_step": "varies"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "fft_hbm_gbs": fft_gbs, "nvlink": nvlink, "stages": stages, "cpu_baseline": cpu,
        }))
    par.free()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
