import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import colore_b200 as cb
from oracle.oracle import tables_from_dump
from test_gpu_parity import _dev_view
g = dict(np.load(os.path.join(ROOT, "tests/golden/ref_n32_lognormal.npz"))); t = tables_from_dump(g)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nc = n // 2 + 1
res = {}
for fused in (1, 0):
    par = cb.ParamCoLoRe(t, n)
    par.set_option("fft_fused", fused)
    gd = _dev_view(par, cb.GRID_DENS)
    gen = torch.Generator(device="cuda").manual_seed(n)
    gd.normal_(generator=gen)
    print("fused", fused, "pitch", par.grid_pitch(), "input std", float(gd.std()), "shape", tuple(gd.shape))
    inp = torch.view_as_complex(gd[:, :, :2 * nc].reshape(n, n, nc, 2).contiguous()).clone()
    cb.fftw_wrap_c2r(par, cb.GRID_DENS); par.synchronize()
    res[fused] = gd[:, :, :n].clone()
    print("  out rms", float(res[fused].double().pow(2).mean().sqrt()), "expected", (2 * n ** 3) ** 0.5)
    par.free()
ref = torch.fft.ifft(inp, dim=0, norm="forward")
ref = torch.fft.ifft(ref, dim=1, norm="forward")
ref[:, :, 0].imag.zero_(); ref[:, :, nc - 1].imag.zero_()
r = torch.fft.irfft(ref, n=n, dim=2, norm="forward")
print("ref rms", float(r.double().pow(2).mean().sqrt()))
for f in (1, 0):
    d = (res[f] - r).abs()
    print("fused", f, "max err", float(d.max()), "frac bad", float((d > 1e-3 * float(r.std())).float().mean()))
    bad = (d > 1e-3 * float(r.std()))
    if bad.any():
        idx = bad.nonzero()
        print("   bad z range", int(idx[:, 0].min()), int(idx[:, 0].max()), "y", int(idx[:, 1].min()), int(idx[:, 1].max()), "x", int(idx[:, 2].min()), int(idx[:, 2].max()))
        print("   bad per z (first 10 planes):", [int(bad[z].sum()) for z in range(10)])
        print("   bad per x-col mod 64 histogram:", torch.bincount(idx[:, 2] % 64, minlength=64).tolist())
# fields: fused vs separate
t = dict(t); t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n))); t["pos_obs"] = 0.5 * t["l_box"]
par = cb.ParamCoLoRe(t, n, seed=5)
cb.create_cartesian_fields(par)
a_d = _dev_view(par, cb.GRID_DENS)[:, :, :n].clone(); a_p = _dev_view(par, cb.GRID_NPOT)[:, :, :n].clone()
par.set_option("fft_fused", 0); par.set_option("fill_fused", 0)
cb.create_cartesian_fields(par); par.synchronize()
b_d = _dev_view(par, cb.GRID_DENS)[:, :, :n]; b_p = _dev_view(par, cb.GRID_NPOT)[:, :, :n]
print("dens: fused std", float(a_d.std()), "sep std", float(b_d.std()), "maxdiff", float((a_d - b_d).abs().max()))
print("npot: fused std", float(a_p.std()), "sep std", float(b_p.std()), "maxdiff", float((a_p - b_p).abs().max()))
d = (a_p - b_p).abs(); bad = d > 1e-3 * float(b_p.std())
print("npot frac bad", float(bad.float().mean()))
if bad.any():
    idx = bad.nonzero()
    print("   bad z", int(idx[:, 0].min()), int(idx[:, 0].max()), "y", int(idx[:, 1].min()), int(idx[:, 1].max()), "x", int(idx[:, 2].min()), int(idx[:, 2].max()))
    print("   a_p absmax", float(a_p.abs().max()), "b_p absmax", float(b_p.abs().max()))
