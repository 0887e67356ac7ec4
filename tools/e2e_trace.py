"""Per-call wall times of the end-to-end loop (synchronous vs asynchronous catalogue read-back): shows which host
call waits when a device-to-host copy queues behind the catalogue copy. python tools/e2e_trace.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import colore_b200 as cb
from bench import build_tables, make_config
n = 1024
cfg = make_config(n); t = build_tables(cfg)
par = cb.ParamCoLoRe(t, n, seed=1); cb.dist.init_comm(par, 0, 1)
par.set_srcs(0, t["srcs_nz_0"], t["srcs_bz_0"])
def step(seed):
    par.seed = seed
    ts = [time.perf_counter()]
    cb.create_cartesian_fields(par); ts.append(time.perf_counter())
    cb.compute_physical_density_field(par); ts.append(time.perf_counter())
    cb.compute_density_normalization(par); ts.append(time.perf_counter())
    k = cb.srcs_set_cartesian(par)[0]; ts.append(time.perf_counter())
    return k, np.diff(ts) * 1e3
for w in range(3): k, _ = step(w)
par.synchronize()
pin = [torch.empty((int(k * 1.2), 9), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
for mode in (0, 1):
    par.set_option("async_results", mode)
    par.synchronize(); t0 = time.perf_counter()
    for s in range(6):
        k, dt = step(100 + s)
        ta = time.perf_counter()
        cb.srcs_get_local_properties(par, 0, out=pin[s & 1][:k])
        tb = time.perf_counter()
        print("mode", mode, "step", s, "create/dens/norm/srcs ms", np.round(dt, 2), "get", round((tb - ta) * 1e3, 2), flush=True)
    par.synchronize()
    print("mode", mode, "total per step", (time.perf_counter() - t0) * 1e3 / 6, flush=True)
par.free()
