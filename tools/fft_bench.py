#!/usr/bin/env python
"""Per-pass timing of the 3-D c2r / r2c transforms alone (fftw_wrap_c2r / fftw_wrap_r2c, fourier.c:81-125).

    python tools/fft_bench.py --n-grid 1024 2048 [--reps 5]
    COLORE_B200_LIB=colore_b200/libcolore_b200_v2.so python tools/fft_bench.py ...   # kernel-variant build

Prints one JSON line per grid size: ms per pass (z, y, x) and achieved GB/s against the algorithmic
8 B/cell/pass.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-grid", type=int, nargs="+", default=[1024])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import colore_b200 as cb
    from bench import build_tables, make_config
    for n in args.n_grid:
        cfg = make_config(n)
        t = build_tables(cfg)
        par = cb.ParamCoLoRe(t, n, seed=3)
        cb.dist.init_comm(par, 0, 1)
        cb.fill_modes(par)
        out = {"n_grid": n, "lib": os.path.basename(cb._lib.SO_PATH)}
        gb = 8.0 * n * n * (n // 2 + 1) * 2 / 1e9          # read + write of one pass
        for name, fn in (("c2r", cb.fftw_wrap_c2r), ("r2c", cb.fftw_wrap_r2c)):
            fn(par, cb.GRID_DENS)
            par.set_profiling(True)
            for _ in range(args.reps):
                fn(par, cb.GRID_NPOT)
            res = {}
            for st in ("fft_z", "fft_y", "fft_x"):
                ms, nl = par.stage_ms(st)
                res[st] = {"ms": ms / nl, "GBps": gb / (ms / nl) * 1e3}
            par.set_profiling(False)
            tot = sum(v["ms"] for v in res.values())
            res["total_ms"] = tot
            res["GBps"] = 3 * gb / tot * 1e3
            out[name] = res
        print(json.dumps(out), flush=True)
        par.free()


if __name__ == "__main__":
    main()
