import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import colore_b200 as cb
from oracle.oracle import RNG_PHILOX, Oracle, tables_from_dump
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 64
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_n32_lognormal.npz")))
t = tables_from_dump(g)
t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
t["pos_obs"] = 0.5 * t["l_box"]
o = Oracle(t, n)
dk, pk = o.fill_modes(RNG_PHILOX, 99)
d0, p0 = o.c2r(dk), o.c2r(pk)
o.normalize_fields(d0, p0)
nzl, iz0 = cb.dist.slab_bounds(n, world, rank)
sl = slice(iz0, iz0 + nzl)
print(rank, "gauss std", d0[:, :, :n].std(), flush=True)
for order, interp in ((1, 1), (2, 2), (2, 0)):
    ref = d0.copy(); o.lpt(ref, order, interp)
    pl = cb.ParamCoLoRe(t, n, dens_type=order, seed=1, nz_here=nzl, iz0_here=iz0, device=local)
    cb.dist.init_comm(pl, rank, world)
    pl.set_option("lpt_interp_type", interp)
    pl.set_option("keep_particles", 1)
    pl.grid_put(cb.GRID_DENS, np.ascontiguousarray(d0[sl]))
    chk = pl.grid_get(cb.GRID_DENS)
    print(rank, "put/get std", chk[:, :, :n].std(), flush=True)
    cb.compute_physical_density_field(pl)
    got = pl.grid_get(cb.GRID_DENS)[:, :, :n].astype(np.float64)
    want = ref[sl, :, :n].astype(np.float64)
    x, y, z = cb.lpt_get_particles(pl)
    print(rank, order, interp, "got std", got.std(), "want std", want.std(), "sum", got.sum(), "maxdiff", np.abs(got - want).max(),
          "exch", cb.lpt_exchange_counts(pl), "z range", z.min(), z.max(), flush=True)
    pl.free()
dist.destroy_process_group()
