"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against
  (1) the golden fixtures written by the UNMODIFIED reference (tests/golden/*.npz), and
  (2) the CPU oracle (oracle/colore_oracle.c) on the same seeded inputs.

Bar (north star): integer / index outputs (per-cell counts, pixel ids, nadd) bit-exact given
identical uniform draws; floating-point fields within the fp32 tolerances written next to each
assert.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import colore_b200 as cb  # noqa: E402
from colore_b200._lib import check  # noqa: E402
from oracle.oracle import RNG_MT, RNG_PHILOX, Oracle, tables_from_dump  # noqa: E402

# ref_n32_bias1 / ref_n32_bias3: the reference compiled with the other bias models (common.h:414-431);
# ref_n32_nosmooth: do_smoothing = 0, smooth_potential = false on a power-of-two grid (fourier.c:347-351 skipped)
# ref_n48_nosmooth: a grid that is not a power of two (mixed-radix transforms, division-based cell indexing)
GOLDEN = ["ref_n32_lognormal", "ref_n32_clip", "ref_n32_bias1", "ref_n32_bias3", "ref_n32_nosmooth", "ref_n48_nosmooth"]


def _bias_model(name):
    return 1 if "bias1" in name else (3 if "bias3" in name else 2)


def _load(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    t = tables_from_dump(g)
    return g, t


def _par(t, **kw):
    return cb.ParamCoLoRe(t, int(t["n_grid"]), dens_type=int(t["dens_type"]), seed=int(t["seed"]),
                          nside_base=int(t["nside_base"]), **kw)


@pytest.fixture(scope="module", params=GOLDEN)
def case(request, golden_dir):
    g, t = _load(golden_dir, request.param)
    bm = _bias_model(request.param)
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]), bias_model=bm)
    par = _par(t, bias_model=bm)
    yield g, t, o, par
    par.free()


def _real(a, n):
    return a[:, :, :n]


# ------------------------------------------------------------------------------------------ FFT
# 24 ... 104: the mixed-radix path (clr_fft_generic.cu): radices 3, 5, 7 and the direct-DFT stage (11, 13)
@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 24, 48, 80, 112, 88, 104])
def test_fft_c2r_r2c_vs_oracle(golden_dir, n):
    g, t = _load(golden_dir, "ref_n32_lognormal")
    o = Oracle(t, n)
    par = cb.ParamCoLoRe(t, n)
    rng = np.random.default_rng(n)
    nc = n // 2 + 1
    ck = (rng.standard_normal((n, n, nc)) + 1j * rng.standard_normal((n, n, nc))).astype(np.complex64)
    par.grid_put(cb.GRID_DENS, ck)
    cb.fftw_wrap_c2r(par, cb.GRID_DENS)
    got = par.grid_get(cb.GRID_DENS)
    ref = o.c2r(ck.copy())
    scale = np.sqrt(np.mean(_real(ref, n).astype(np.float64) ** 2))
    err = np.abs(_real(got, n) - _real(ref, n)).max() / scale
    assert err < 2e-5, f"c2r n={n}: max error {err:.2e} of rms (fp32 tolerance 2e-5)"
    # forward transform of a real field
    x = np.zeros((n, n, 2 * nc), np.float32)
    x[:, :, :n] = rng.standard_normal((n, n, n)).astype(np.float32)
    par.grid_put(cb.GRID_NPOT, x)
    cb.fftw_wrap_r2c(par, cb.GRID_NPOT)
    gotk = par.grid_get(cb.GRID_NPOT).view(np.complex64)
    refk = o.r2c(x.copy())
    errk = np.abs(gotk - refk).max() / np.sqrt(np.mean(np.abs(refk) ** 2))
    assert errk < 2e-5, f"r2c n={n}: max error {errk:.2e} of rms"
    # round trip c2r(r2c(x)) = n^3 x  (size-independent property)
    cb.fftw_wrap_c2r(par, cb.GRID_NPOT)
    back = par.grid_get(cb.GRID_NPOT)
    assert np.abs(_real(back, n) / n ** 3 - _real(x, n)).max() < 2e-5
    par.free()


def test_fft_drops_imag_of_xdc_and_nyquist(golden_dir):
    """FFTW c2r semantics on non-Hermitian input."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    n, nc = 32, 17
    par = cb.ParamCoLoRe(t, n)
    rng = np.random.default_rng(5)
    ck = (rng.standard_normal((n, n, nc)) + 1j * rng.standard_normal((n, n, nc))).astype(np.complex64)
    par.grid_put(cb.GRID_DENS, ck)
    cb.fftw_wrap_c2r(par, cb.GRID_DENS)
    a = par.grid_get(cb.GRID_DENS)[:, :, :n].copy()
    # numpy's irfftn follows the same convention (complex axes first, real axis last)
    ref = np.fft.irfftn(ck.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * n ** 3
    assert np.abs(a - ref).max() / ref.std() < 2e-5
    par.free()


@pytest.mark.parametrize("n", [300, 360])
def test_fft_mixed_radix_vs_numpy(golden_dir, n):
    """Grids that are not powers of two at a size where several stages of every radix run (300 = 4 * 3 * 5 * 5,
    360 = 4 * 2 * 3 * 3 * 5), against numpy / pocketfft in double precision."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    par = cb.ParamCoLoRe(t, n)
    rng = np.random.default_rng(n)
    nc = n // 2 + 1
    ck = (rng.standard_normal((n, n, nc)) + 1j * rng.standard_normal((n, n, nc))).astype(np.complex64)
    par.grid_put(cb.GRID_DENS, ck)
    cb.fftw_wrap_c2r(par, cb.GRID_DENS)
    got = par.grid_get(cb.GRID_DENS)[:, :, :n]
    ref = np.fft.irfftn(ck.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * float(n) ** 3
    assert np.abs(got - ref).max() / ref.std() < 2e-5
    x = np.zeros((n, n, 2 * nc), np.float32)
    x[:, :, :n] = rng.standard_normal((n, n, n)).astype(np.float32)
    par.grid_put(cb.GRID_NPOT, x)
    cb.fftw_wrap_r2c(par, cb.GRID_NPOT)
    gotk = par.grid_get(cb.GRID_NPOT).view(np.complex64)
    refk = np.fft.rfftn(x[:, :, :n].astype(np.float64), axes=(0, 1, 2))
    assert np.abs(gotk - refk).max() / np.sqrt(np.mean(np.abs(refk) ** 2)) < 2e-5
    par.free()


class _DevArray:
    """Raw device pointer -> torch, through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def _dev_view(par, which):
    import torch
    pitch = par.grid_pitch()
    t = torch.as_tensor(_DevArray(par.grid_device_ptr(which), par.nz_here * par.n_grid * pitch), device="cuda:0")
    return t.view(par.nz_here, par.n_grid, pitch)


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("n", [512, 1024, 2048])
def test_fft_large_vs_cufft(golden_dir, n, fused):
    """The transforms the bench times (FftPlan<512/1024/2048>, the fused y+x pass and the three separate passes) against
    an INDEPENDENT implementation: cuFFT through torch.fft, test-only, applied axis by axis in
    chunks (complex passes over z and y, then half-complex -> real over x with Im of the x-DC / x-Nyquist lines
    dropped = FFTW's c2r on a non-Hermitian spectrum). Tolerance: 2e-5 of the rms of the result, L-infinity."""
    import torch
    if n == 2048 and torch.cuda.mem_get_info()[1] < 150e9:
        pytest.skip("needs ~125 GB of device memory")
    g, t = _load(golden_dir, "ref_n32_lognormal")
    par = cb.ParamCoLoRe(t, n)
    par.set_option("fft_fused", fused)
    nc = n // 2 + 1
    gd = _dev_view(par, cb.GRID_DENS)
    gen = torch.Generator(device="cuda").manual_seed(n)
    # ---- c2r: random non-Hermitian half spectrum, written straight into the device grid
    for z0 in range(0, n, 64):
        gd[z0:z0 + 64].normal_(generator=gen)
    ref = torch.empty((n, n, nc), dtype=torch.complex64, device="cuda")
    for z0 in range(0, n, 64):
        ref[z0:z0 + 64] = torch.view_as_complex(gd[z0:z0 + 64, :, :2 * nc].reshape(-1, n, nc, 2).contiguous())
    torch.cuda.synchronize()         # the library runs on its own (non-blocking) stream: torch's writes must have landed
    cb.fftw_wrap_c2r(par, cb.GRID_DENS)
    par.synchronize()
    cy = max(1, (1 << 27) // (n * n))                      # lines per chunk: ~1 GB temporaries
    for x0 in range(0, nc, cy):
        ref[:, :, x0:x0 + cy] = torch.fft.ifft(ref[:, :, x0:x0 + cy], dim=0, norm="forward")
    for z0 in range(0, n, cy):
        ref[z0:z0 + cy] = torch.fft.ifft(ref[z0:z0 + cy], dim=1, norm="forward")
    ref[:, :, 0].imag.zero_()
    ref[:, :, nc - 1].imag.zero_()
    err, s2 = 0.0, 0.0
    for z0 in range(0, n, cy):
        r = torch.fft.irfft(ref[z0:z0 + cy], n=n, dim=2, norm="forward")
        err = max(err, float((gd[z0:z0 + cy, :, :n] - r).abs().max()))
        s2 += float((r.double() ** 2).sum())
    rms = (s2 / float(n) ** 3) ** 0.5
    assert err < 2e-5 * rms, f"c2r n={n} fused={fused}: max error {err / rms:.2e} of rms (fp32 tolerance 2e-5)"
    # ---- r2c of a real field against rfft / fft / fft
    for z0 in range(0, n, 64):
        gd[z0:z0 + 64].normal_(generator=gen)
    for z0 in range(0, n, cy):
        ref[z0:z0 + cy] = torch.fft.rfft(gd[z0:z0 + cy, :, :n], dim=2)
    torch.cuda.synchronize()
    cb.fftw_wrap_r2c(par, cb.GRID_DENS)
    par.synchronize()
    for z0 in range(0, n, cy):
        ref[z0:z0 + cy] = torch.fft.fft(ref[z0:z0 + cy], dim=1)
    err, s2 = 0.0, 0.0
    for x0 in range(0, nc, cy):
        r = torch.fft.fft(ref[:, :, x0:x0 + cy], dim=0)
        got = torch.view_as_complex(gd[:, :, 2 * x0:2 * min(nc, x0 + cy)].reshape(n, n, -1, 2).contiguous())
        err = max(err, float((got - r).abs().max()))
        s2 += float((r.abs().double() ** 2).sum())
    rms = (s2 / (float(n) ** 2 * nc)) ** 0.5
    assert err < 2e-5 * rms, f"r2c n={n}: max error {err / rms:.2e} of rms"
    del ref
    torch.cuda.empty_cache()
    par.free()


@pytest.mark.parametrize("n,fill_w,cluster", [(128, 8, 0), (256, 8, 0), (256, 4, 0), (256, 8, 1)])
def test_fused_fields_vs_oracle(golden_dir, n, fill_w, cluster):
    """n >= 128 on one GPU: the mode fill is fused into the z pass and the y + x passes run as one kernel through L2
    (clr_fft.cu: fill_z_kernel, yx_fused_kernel). Same Philox stream in the oracle -> fields to 2e-5 sigma.
    fill_w = 4: half-width tiles (optional at 1024^3); cluster: one field per CTA of a pair, the fill shared through
    distributed shared memory (fill_z_cluster_kernel, what runs at 2048^3)."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    t = dict(t)
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    o = Oracle(t, n)
    par = cb.ParamCoLoRe(t, n, seed=77)
    par.set_option("fill_w", fill_w)
    par.set_option("fill_cluster", cluster)
    dk, pk = o.fill_modes(RNG_PHILOX, 77)
    dens, npot = o.c2r(dk), o.c2r(pk)
    o.normalize_fields(dens, npot)
    _, s2_ref = o.sigma_dens(dens)
    mean, s2 = cb.create_cartesian_fields(par)
    got_d, got_p = par.grid_get(cb.GRID_DENS), par.grid_get(cb.GRID_NPOT)
    assert np.abs(_real(got_d, n) - _real(dens, n)).max() < 2e-5 * np.sqrt(s2_ref)
    assert np.abs(_real(got_p, n) - _real(npot, n)).max() < 2e-5 * _real(npot, n).std()
    assert abs(s2 / s2_ref - 1) < 1e-5 and abs(mean) < 1e-5 * np.sqrt(s2_ref)
    par.free()


@pytest.mark.parametrize("n,fill_w,cluster", [(128, 8, 0), (512, 8, 0), (1024, 8, 0), (1024, 4, 0), (1024, 8, 1), (2048, 8, 1),
                                              (2048, 4, 0)])
def test_fused_fields_match_separate_passes(golden_dir, n, fill_w, cluster):
    """Full-size consistency of the two code paths of create_cartesian_fields: fused (fill + z pass, y + x pass) against
    stand-alone fill + three separate axis passes, same seed. Both evaluate the same butterflies in fp32.
    At 2048^3 (two 34 GB grids + scratch) every 16th plane is kept for the comparison, plus the sum of every plane."""
    import torch
    if n >= 2048 and torch.cuda.mem_get_info()[1] < 150e9:
        pytest.skip("needs ~150 GB of device memory")
    g, t = _load(golden_dir, "ref_n32_lognormal")
    t = dict(t)
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    par = cb.ParamCoLoRe(t, n, seed=5)
    par.set_option("fill_w", fill_w)
    par.set_option("fill_cluster", cluster)
    zs = 16 if n >= 2048 else 1

    def snapshot(grid):
        v = _dev_view(par, grid)
        sums = torch.stack([v[z0:z0 + 64, :, :n].sum(dim=(1, 2), dtype=torch.float64) for z0 in range(0, n, 64)]).flatten()
        return v[::zs, :, :n].clone(), sums

    mean1, s2_1 = cb.create_cartesian_fields(par)
    par.synchronize()
    a_d, a_ds = snapshot(cb.GRID_DENS)
    a_p, a_ps = snapshot(cb.GRID_NPOT)
    torch.cuda.synchronize()         # the library runs on its own (non-blocking) stream
    par.set_option("fft_fused", 0)
    par.set_option("fill_fused", 0)
    mean2, s2_2 = cb.create_cartesian_fields(par)
    par.synchronize()
    b_d, b_p = _dev_view(par, cb.GRID_DENS)[::zs, :, :n], _dev_view(par, cb.GRID_NPOT)[::zs, :, :n]
    sig_p = float(b_p[:: max(1, 64 // zs)].std())
    assert float((a_d - b_d).abs().max()) < 5e-6 * np.sqrt(s2_2)
    assert float((a_p - b_p).abs().max()) < 5e-6 * sig_p
    assert abs(s2_1 / s2_2 - 1) < 1e-6
    del a_d, a_p, b_d, b_p
    torch.cuda.empty_cache()
    # plane sums: sum over n^2 cells of errors <= 5e-6 sigma each
    _, b_ds = snapshot(cb.GRID_DENS)
    _, b_ps = snapshot(cb.GRID_NPOT)
    assert float((a_ds - b_ds).abs().max()) < 5e-6 * np.sqrt(s2_2) * n * n
    assert float((a_ps - b_ps).abs().max()) < 5e-6 * sig_p * n * n
    torch.cuda.empty_cache()
    par.free()


# ------------------------------------------------------------------------------ Gaussian fields
@pytest.mark.parametrize("exact", [1, 0])
def test_fill_modes_vs_oracle(case, exact):
    g, t, o, par = case
    dk, pk = o.fill_modes(RNG_PHILOX, int(t["seed"]))
    par.set_option("exact_math", exact)
    cb.fill_modes(par)
    par.set_option("exact_math", 0)
    gd = par.grid_get(cb.GRID_DENS).view(np.complex64)
    gp = par.grid_get(cb.GRID_NPOT).view(np.complex64)
    for got, ref in ((gd, dk), (gp, pk)):
        err = np.abs(got - ref)
        big = np.abs(ref) > 1e-3 * np.abs(ref).max()
        if exact:
            # double arithmetic on both sides: agreement to a few fp32 ulps of each mode
            assert (err / np.maximum(np.abs(ref), 1e-30))[np.abs(ref) > 0].max() < 5e-7
        else:
            # fp32 transcendentals: 2e-6 of each significant mode, 1e-6 of the largest mode overall
            assert (err[big] / np.abs(ref[big])).max() < 2e-6
            assert err.max() < 1e-6 * np.abs(ref).max()
        assert np.all(got[np.abs(ref) == 0] == 0)


def test_injected_white_noise_matches_reference(case):
    """Reference's own MT19937 modes injected -> GPU FFT + scaling + sigma vs the reference fields."""
    g, t, o, par = case
    n = o.n
    dk, pk = o.fill_modes(RNG_MT, int(t["seed"]))
    par.grid_put(cb.GRID_DENS, dk)
    par.grid_put(cb.GRID_NPOT, pk)
    mean, s2 = cb.create_cartesian_fields(par, inject=True)
    dens = par.grid_get(cb.GRID_DENS)
    npot = par.grid_get(cb.GRID_NPOT)
    sig = np.sqrt(g["s1_sigma2_gauss"][0])
    assert np.abs(_real(dens, n) - _real(g["s1_dens_gauss"], n)).max() < 2e-5 * sig
    assert np.abs(_real(npot, n) - _real(g["s1_npot"], n)).max() < 2e-5 * _real(g["s1_npot"], n).std()
    assert abs(s2 / g["s1_sigma2_gauss"][0] - 1) < 1e-5
    assert abs(mean) < 1e-5 * sig


def test_own_stream_end_to_end_vs_oracle(case):
    """GPU counter-based stream, whole create_cartesian_fields, against the oracle's Philox run."""
    g, t, o, par = case
    n = o.n
    dk, pk = o.fill_modes(RNG_PHILOX, int(t["seed"]))
    dens, npot = o.c2r(dk), o.c2r(pk)
    o.normalize_fields(dens, npot)
    _, s2_ref = o.sigma_dens(dens)
    mean, s2 = cb.create_cartesian_fields(par)
    got = par.grid_get(cb.GRID_DENS)
    assert np.abs(_real(got, n) - _real(dens, n)).max() < 2e-5 * np.sqrt(s2_ref)
    assert abs(s2 / s2_ref - 1) < 1e-5


# --------------------------------------------------------------------------- physical density
@pytest.mark.parametrize("exact", [1, 0])
def test_physical_density_vs_reference(case, exact):
    g, t, o, par = case
    n = o.n
    par.grid_put(cb.GRID_DENS, g["s1_dens_gauss"])
    par.set_sigma2_gauss(g["s1_sigma2_gauss"][0])
    par.set_option("exact_math", exact)
    cb.compute_physical_density_field(par)
    par.set_option("exact_math", 0)
    got = _real(par.grid_get(cb.GRID_DENS), n).astype(np.float64)
    ref = _real(g["s2_dens"], n).astype(np.float64)
    if exact:
        # same double arithmetic, CUDA exp vs glibc exp: at most 1 fp32 ulp apart
        assert np.abs(got - ref).max() <= 2.5e-7 * (1 + np.abs(ref).max())
        assert np.mean(got == ref) > 0.99
    else:
        # fp32 evaluation: 1e-6 relative on 1+delta (the physical density), the stated fp32 tolerance
        assert (np.abs(got - ref) / (1 + np.abs(ref))).max() < 1e-6


@pytest.mark.parametrize("exact", [1, 0])
def test_density_normalization_vs_reference(case, exact):
    g, t, o, par = case
    rtol = 1e-12 if exact else 1e-6      # fp32 bias_model terms summed in double: ~1e-7 systematic
    par.set_option("exact_math", exact)
    par.grid_put(cb.GRID_DENS, g["s2_dens"])
    npop = sum(1 for k in t if k.startswith("srcs_bz_"))
    for i in range(npop):
        par.set_srcs(i, t[f"srcs_nz_{i}"], t[f"srcs_bz_{i}"])
    if "imap_bz_0" in t:
        par.set_imap(0, t["imap_tz_0"], t["imap_bz_0"], 8, g["s4_imap_r0_0"], g["s4_imap_rf_0"])
    cb.compute_density_normalization(par)
    par.set_option("exact_math", 0)
    for i in range(npop):
        norm, ends, zends = cb.get_norm(par, 0, i)
        np.testing.assert_allclose(norm, g[f"s3_srcs_norm_{i}"], rtol=rtol)
        np.testing.assert_allclose(ends, g[f"s3_srcs_norm_ends_{i}"], rtol=rtol)
        np.testing.assert_allclose(zends, g["s3_znorm_ends"], rtol=1e-13, atol=0)
    if "imap_bz_0" in t:
        norm, ends, _ = cb.get_norm(par, 1, 0)
        np.testing.assert_allclose(norm, g["s3_imap_norm_0"], rtol=rtol)


def test_fused_lognormal_histogram(case):
    """Default flow with the populations known before the density transform: lognormalize / densclip and the
    normalisation histogram run as ONE pass (clr_fields.cu: norm_hist_fast_kernel<.., XFORM>). Both the field and the
    normalisation tables must match the reference (fp32 tolerances) and the separate passes."""
    g, t, o, par = case
    n = o.n
    npop = sum(1 for k in t if k.startswith("srcs_bz_"))
    for i in range(npop):
        par.set_srcs(i, t[f"srcs_nz_{i}"], t[f"srcs_bz_{i}"])
    res = {}
    for fused in (1, 0):
        par.set_option("hist_fused", fused)
        par.grid_put(cb.GRID_DENS, g["s1_dens_gauss"])
        par.set_sigma2_gauss(g["s1_sigma2_gauss"][0])
        cb.compute_physical_density_field(par)
        cb.compute_density_normalization(par)
        res[fused] = (_real(par.grid_get(cb.GRID_DENS), n).astype(np.float64), [cb.get_norm(par, 0, i)[0] for i in range(npop)])
    par.set_option("hist_fused", 1)
    ref = _real(g["s2_dens"], n).astype(np.float64)
    assert (np.abs(res[1][0] - ref) / (1 + np.abs(ref))).max() < 1e-6
    assert (np.abs(res[1][0] - res[0][0]) / (1 + np.abs(ref))).max() < 5e-7
    for i in range(npop):
        np.testing.assert_allclose(res[1][1][i], g[f"s3_srcs_norm_{i}"], rtol=1e-6)
        np.testing.assert_allclose(res[1][1][i], res[0][1][i], rtol=1e-6)


# ------------------------------------------------------------------------------------ sources
def _setup_sources(g, t, par):
    par.grid_put(cb.GRID_DENS, g["s2_dens"])
    par.grid_put(cb.GRID_NPOT, g["s1_npot"])
    par.update_halo()
    npop = sum(1 for k in t if k.startswith("srcs_bz_"))
    for i in range(npop):
        par.set_srcs(i, t[f"srcs_nz_{i}"], t[f"srcs_bz_{i}"])
        cb.set_norm(par, 0, i, g[f"s3_srcs_norm_{i}"], g[f"s3_srcs_norm_ends_{i}"])
    return npop


def test_sources_bit_exact_vs_oracle(case):
    g, t, o, par = case
    npop = _setup_sources(g, t, par)
    o.set_halo(g["s1_npot"])
    seed = int(t["seed"])
    nsrc = cb.srcs_set_cartesian(par)
    for ipop in range(npop):
        ends = g[f"s3_srcs_norm_ends_{ipop}"]
        ns, tot = o.srcs_poisson(g["s2_dens"], t[f"srcs_nz_{ipop}"], t[f"srcs_bz_{ipop}"], g[f"s3_srcs_norm_{ipop}"],
                                 ends[0], ends[1], RNG_PHILOX, seed, ipop)
        assert nsrc[ipop] == tot
        assert np.array_equal(cb.srcs_get_counts(par, ipop), ns), "per-cell Poisson counts differ"
        pos_ref, ipix_ref = o.srcs_place(g["s1_npot"], ns, RNG_PHILOX, seed, ipop)
        pos, ipix = cb.srcs_get_cartesian(par, ipop)
        assert np.array_equal(ipix, ipix_ref), "base-pixel indices differ"
        assert np.array_equal(pos[:, :3], pos_ref[:, :3]), "in-cell positions differ"
        np.testing.assert_allclose(pos[:, 3], pos_ref[:, 3], rtol=2e-6, atol=1e-12)
        srcs = cb.srcs_get_local_properties(par, ipop)
        ref = o.srcs_local_properties(pos_ref)
        np.testing.assert_allclose(srcs[:, :3], ref[:, :3], rtol=3e-7, atol=3e-5)   # ra, dec [deg], z
        assert np.array_equal(srcs[:, 4:6], ref[:, 4:6])
        # statistical sanity against the reference's own MT19937 catalogue: same expected number
        n_ref = g[f"s4_srcs_ipix_{ipop}"].size
        assert abs(tot - n_ref) < 6 * np.sqrt(n_ref)


@pytest.mark.parametrize("scale", [1.0, 0.15, 0.04])
def test_poisson_dense_bit_exact(golden_dir, scale):
    """gsl_ran_poisson's mu > 10 branch (gamma / binomial reduction, common.c:187; dev_gamma_int, dev_gamma_large,
    dev_binomial in clr_srcs.cu) against the oracle, bit for bit. ref_n32_dense holds ~120 sources per cell; the scale
    factors put the bulk of the occupied cells at lambda ~ 80-1000, ~12-150 and ~3-40 (the switch at mu = 10, where the
    fp32 screen must hand over to the exact path)."""
    g, t = _load(golden_dir, "ref_n32_dense")
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]))
    par = _par(t)
    nz = t["srcs_nz_0"] * scale
    par.grid_put(cb.GRID_DENS, g["s2_dens"])
    par.grid_put(cb.GRID_NPOT, g["s1_npot"])
    par.update_halo()
    par.set_srcs(0, nz, t["srcs_bz_0"])
    ends = g["s3_srcs_norm_ends_0"]
    cb.set_norm(par, 0, 0, g["s3_srcs_norm_0"], ends)
    seed = int(t["seed"])
    nsrc = cb.srcs_set_cartesian(par)[0]
    ns, tot = o.srcs_poisson(g["s2_dens"], nz, t["srcs_bz_0"], g["s3_srcs_norm_0"], ends[0], ends[1], RNG_PHILOX, seed, 0)
    got = cb.srcs_get_counts(par, 0)
    assert (ns > 10).mean() > 0.1 and (scale > 0.1 or ((ns > 0) & (ns <= 10)).mean() > 0.02)
    assert nsrc == tot and np.array_equal(got, ns), f"{(got != ns).sum()} cells differ"
    o.set_halo(g["s1_npot"])
    pos_ref, ipix_ref = o.srcs_place(g["s1_npot"], ns, RNG_PHILOX, seed, 0)
    pos, ipix = cb.srcs_get_cartesian(par, 0)
    assert np.array_equal(ipix, ipix_ref) and np.array_equal(pos[:, :3], pos_ref[:, :3])
    par.free()


def test_async_results_match_sync(case):
    """Option async_results: the catalogue read-back overlaps the next run and returns the same records."""
    import torch
    g, t, o, par = case
    _setup_sources(g, t, par)
    cb.srcs_set_cartesian(par)
    want = cb.srcs_get_local_properties(par, 0).copy()
    pinned = torch.empty(want.shape, dtype=torch.float32).pin_memory()
    out = pinned.numpy()
    out[:] = -1
    par.set_option("async_results", 1)
    cb.srcs_get_local_properties(par, 0, out=out)
    cb.srcs_set_cartesian(par)          # the next run must wait for the copy before it reuses the buffers
    par.synchronize()
    par.set_option("async_results", 0)
    assert np.array_equal(out, want)
    assert np.array_equal(cb.srcs_get_local_properties(par, 0), want)


def test_three_populations_one_of_them_empty(golden_dir, tmp_path):
    """Several populations at once (srcs.c:285-294 loops over n_srcs; density.c:1128-1393 normalises them in ONE walk:
    the NPOP = 3 instantiation of the histogram kernel) and a population without sources (n(z) = 0): normalisation
    tables against the oracle, per-cell counts and pixels bit-exact per population (each on its own Philox streams),
    an empty catalogue through every read-back and through the writer (header only, as io.c writes it)."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]), bias_model=2)
    par = _par(t, bias_model=2)
    try:
        nz0, bz0 = t["srcs_nz_0"], t["srcs_bz_0"]
        pops = [(nz0, bz0), (0.5 * nz0, bz0 + 0.3), (np.zeros_like(nz0), bz0)]
        par.grid_put(cb.GRID_DENS, g["s2_dens"])
        par.grid_put(cb.GRID_NPOT, g["s1_npot"])
        par.update_halo()
        for i, (nz, bz) in enumerate(pops):
            par.set_srcs(i, nz, bz)
        cb.compute_density_normalization(par)
        nm = o.density_normalization(g["s2_dens"].copy(), [bz for _, bz in pops])
        o.set_halo(g["s1_npot"])
        seed = int(t["seed"])
        nsrc = cb.srcs_set_cartesian(par)
        for i, (nz, bz) in enumerate(pops):
            norm, ends, _ = cb.get_norm(par, 0, i)
            np.testing.assert_allclose(norm, nm["norm"][i], rtol=1e-6)
            ns, tot = o.srcs_poisson(g["s2_dens"], nz, bz, norm, ends[0], ends[1], RNG_PHILOX, seed, i)
            assert nsrc[i] == tot and np.array_equal(cb.srcs_get_counts(par, i), ns)
            pos_ref, ipix_ref = o.srcs_place(g["s1_npot"], ns, RNG_PHILOX, seed, i)
            pos, ipix = cb.srcs_get_cartesian(par, i)
            assert np.array_equal(ipix, ipix_ref) and np.array_equal(pos[:, :3], pos_ref[:, :3])
        assert nsrc[0] > 0 and 0 < nsrc[1] < nsrc[0] and nsrc[2] == 0
        assert cb.srcs_get_local_properties(par, 2).shape == (0, 9)
        cb.srcs_beams(par)                                  # no-op on the empty catalogue
        fa = str(tmp_path / "empty.txt")
        cb.write_catalog(par, 2, fa, "ascii")
        assert open(fa).read() == "#[1]type [2]RA, [3]dec, [4]z0, [5]dz_RSD \n"
        ff = str(tmp_path / "empty.fits")
        cb.write_catalog(par, 2, ff, "fits")
        raw = open(ff, "rb").read()
        assert len(raw) == 5760 and b"NAXIS2  =                    0" in raw
    finally:
        par.free()


def test_bad_arguments_fail_loudly(golden_dir):
    """Error behaviour of the boundary: every ABI call returns non-zero + clr_last_error (mapped to report_error(1, ...)
    by the glue, common.c:290-306) instead of computing something else."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    for n in (30, 4 * 37):                                  # not a multiple of 4; a prime factor above 31
        with pytest.raises(cb.ColoreError):
            cb.ParamCoLoRe(t, n)
    par = _par(t)
    try:
        with pytest.raises(cb.ColoreError, match="unknown option"):
            par.set_option("no_such_option", 1)
        with pytest.raises(cb.ColoreError, match="out of range"):
            check(par.lib.clr_srcs_distribute(par.ctx, 99, 0, None))          # population index
        par.set_srcs(0, t["srcs_nz_0"], t["srcs_bz_0"])
        with pytest.raises(cb.ColoreError, match="open file"):
            cb.write_catalog(par, 0, "/nonexistent_dir/x.txt", "ascii")        # common.c:58-62 error_open_file
    finally:
        par.free()


def test_beam_rsd_vs_oracle(case):
    g, t, o, par = case
    _setup_sources(g, t, par)
    o.set_halo(g["s1_npot"])
    cb.srcs_set_cartesian(par)
    pos, _ = cb.srcs_get_cartesian(par, 0)
    before = cb.srcs_get_local_properties(par, 0)
    cb.srcs_beams(par)
    after = cb.srcs_get_local_properties(par, 0)
    ref = o.srcs_beam_rsd(g["s1_npot"], pos, before.copy())
    np.testing.assert_allclose(after[:, 3], ref[:, 3], rtol=1e-5, atol=1e-9)
    assert np.all(after[:, 4:6] == 0)


@pytest.mark.parametrize("name", ["ref_n32_lensing", "ref_n32_gskw"])
def test_srcs_lensing_and_skewers_vs_oracle(golden_dir, name):
    """srcs.c:425-744: per-source shear / convergence / deflection from the NGP tidal + velocity stencils
    along every ray, density or Gaussian skewers (CIC) and their post-processing, on the GPU's own catalogue against the
    oracle, which reproduces the unmodified reference bit for bit on these two configurations (CPU suite)."""
    g, t = _load(golden_dir, name)
    flags = g["s6_srcs_flags_0"].astype(int)
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]))
    par = _par(t)
    _setup_sources(g, t, par)
    par.set_sigma2_gauss(g["s1_sigma2_gauss"][0])
    o.set_halo(g["s1_npot"])
    cb.srcs_set_cartesian(par)
    pos, _ = cb.srcs_get_cartesian(par, 0)
    assert pos.shape[0] > 3000
    before = cb.srcs_get_local_properties(par, 0)
    srcs, dg, vs = cb.srcs_get_beam_properties(par, 0, lensing=bool(flags[0]), skewers=bool(flags[1]),
                                               gaussian_skewers=bool(flags[2]))
    ref = before.copy()
    ref[:, 6:] = 0
    ref, rdg, rvs = o.srcs_beam_full(g["s2_dens"], g["s1_npot"], pos, ref, g["s1_sigma2_gauss"][0], lensing=flags[0],
                                     skewers=flags[1], gaussian=flags[2])
    assert np.array_equal(srcs[:, :3], ref[:, :3])                                  # ra, dec, z0 untouched
    np.testing.assert_allclose(srcs[:, 3], ref[:, 3], rtol=1e-5, atol=1e-9)         # dz_rsd
    if flags[0]:
        for col in range(4, 9):                                                     # e1, e2, kappa, dra, ddec
            np.testing.assert_allclose(srcs[:, col], ref[:, col], rtol=2e-5, atol=2e-6 * np.abs(ref[:, col]).max())
        assert np.abs(ref[:, 6]).max() > 0
    else:
        assert np.all(srcs[:, 4:6] == 0)
    if flags[1]:
        # sources beyond r_max - dr/2 exist, so the overrun of srcs.c:726 into the next skewer is exercised
        r = np.sqrt((pos[:, :3].astype(np.float64) ** 2).sum(1))
        nr = o.n // 2
        assert ((r * nr / t["r_max"] + 0.5).astype(int) > nr - 1).any()
        np.testing.assert_allclose(dg, rdg, rtol=2e-6, atol=2e-6 * np.abs(rdg).max())
        np.testing.assert_allclose(vs, rvs, rtol=1e-5, atol=1e-6 * np.abs(rvs).max())
        assert np.abs(rvs).max() > 0 and np.abs(rdg).max() > 0
    par.free()


def test_fast_lensing_shells_and_sources_vs_reference(golden_dir):
    """lensing.c:39-250 + srcs.c:666-723 (reference builds with -D_USE_FAST_LENSING): the adaptive shells
    against the unmodified reference's, then the interpolation onto the GPU's own catalogue against the oracle (which
    reproduces the reference bit for bit, CPU suite) fed with the same shells."""
    g, t = _load(golden_dir, "ref_n32_fastlens")
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]))
    par = _par(t)
    _setup_sources(g, t, par)
    npp, nside_sh = g["s6_lens_npp"], g["s6_lens_nside"]
    nr = len(npp)
    pos_sh = g["s6_lens_pos"].reshape(-1, 3)
    nbeams = pos_sh.shape[0] // int(npp[-1])
    r0 = ((np.arange(nr) + 1) * np.float32(np.float32(t["r_max"]) / np.float32(nr))).astype(np.float32)   # cosmo.c:851-860
    data, r_snap = cb.lensing_get_beam_properties(par, r0, npp, pos_sh)
    assert np.array_equal(r_snap, g["s6_lens_r"])
    ref = np.concatenate([g[f"s6_lens_data_{i:03d}"] for i in range(nr)])
    assert data.shape == ref.shape
    # fine pixels are summed into the coarse shells with float atomics: summation order differs
    np.testing.assert_allclose(data, ref, rtol=2e-5, atol=2e-6 * np.abs(ref).max())
    cb.srcs_set_cartesian(par)
    pos, _ = cb.srcs_get_cartesian(par, 0)
    cb.srcs_beams(par)
    before = cb.srcs_get_local_properties(par, 0)
    srcs, bad = cb.srcs_lensing_from_shells(par, 0, r_snap, nside_sh)
    assert bad == 0 and pos.shape[0] > 3000
    want = before.copy()
    want, bad_o = o.srcs_fast_lensing(r_snap, nside_sh, npp, data, nbeams, pos, want)
    assert bad_o == 0
    assert np.array_equal(srcs[:, :4], want[:, :4])
    np.testing.assert_allclose(srcs[:, 4:], want[:, 4:], rtol=1e-6, atol=1e-7 * np.abs(want[:, 4:]).max())
    assert np.abs(want[:, 6]).max() > 0
    par.free()


def test_custom_map_vs_reference(golden_dir):
    """cstm.c:38-145 against the unmodified reference's map, and the custom population's share of the
    density normalisation (density.c:1177-1178, 1315-1354) against its table."""
    g, t = _load(golden_dir, "ref_n32_lensing")
    par = _par(t)
    par.set_option("exact_math", 1)
    par.grid_put(cb.GRID_DENS, g["s2_dens"])
    par.set_sigma2_gauss(g["s1_sigma2_gauss"][0])
    par.set_srcs(0, t["srcs_nz_0"], t["srcs_bz_0"])
    par.set_cstm(0, t["cstm_kz_0"], t["cstm_bz_0"])
    cb.compute_density_normalization(par)
    norm, ends, _ = cb.get_norm(par, 2, 0)
    np.testing.assert_allclose(norm, g["s3_cstm_norm_0"], rtol=1e-12)
    np.testing.assert_allclose(ends, g["s3_cstm_norm_ends_0"], rtol=1e-12)
    cb.set_norm(par, 2, 0, g["s3_cstm_norm_0"], g["s3_cstm_norm_ends_0"])
    pos = g["s6_cstm_pos_0"].reshape(-1, 3)
    data = cb.cstm_get_beam_properties(par, 0, pos)
    ref = g["s6_cstm_data_0"]
    np.testing.assert_allclose(data, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
    par.free()


# --------------------------------------------------------------------------------------- maps
@pytest.mark.parametrize("exact", [1, 0])
def test_maps_vs_reference(golden_dir, exact):
    g, t = _load(golden_dir, "ref_n32_lognormal")
    par = _par(t)
    par.set_option("exact_math", exact)      # 0: fp32-screened painter (default), 1: the double kernel
    par.grid_put(cb.GRID_DENS, g["s2_dens"])
    par.grid_put(cb.GRID_NPOT, g["s1_npot"])
    par.update_halo()
    r0, rf = g["s4_imap_r0_0"], g["s4_imap_rf_0"]
    nside = int(np.sqrt(g["s4_imap_data_0"].size / len(r0) / 12))
    par.set_imap(0, t["imap_tz_0"], t["imap_bz_0"], nside, r0, rf)
    cb.set_norm(par, 1, 0, g["s3_imap_norm_0"], g["s3_imap_norm_ends_0"])
    data, nadd = cb.imap_set_cartesian(par, 0)
    assert np.array_equal(nadd.ravel(), g["s4_imap_nadd_0"]), "imap hit counts differ"
    ref = g["s4_imap_data_0"].reshape(data.shape)
    # float atomics: summation order differs (it does in the reference under OpenMP too)
    np.testing.assert_allclose(data, ref, rtol=2e-5, atol=1e-7 * np.abs(ref).max())
    # kappa / ISW: per-pixel double accumulators in the same order -> fp32-rounding agreement
    listpix, pos = cb.healpix.hp_shell_pixels(int(np.sqrt(g["s6_kappa_listpix"].size / 12)), int(t["nside_base"]))
    assert np.array_equal(listpix, g["s6_kappa_listpix"])
    assert np.array_equal(pos.ravel(), g["s6_kappa_pos"])
    kap = cb.kappa_get_beam_properties(par, pos, g["s6_kappa_rf"])
    refk = g["s6_kappa_data"].reshape(kap.shape)
    np.testing.assert_allclose(kap, refk, rtol=1e-5, atol=1e-6 * np.abs(refk).max())
    isw = cb.isw_get_beam_properties(par, pos, g["s6_isw_rf"])
    refi = g["s6_isw_data"].reshape(isw.shape)
    np.testing.assert_allclose(isw, refi, rtol=1e-5, atol=1e-6 * np.abs(refi).max())
    par.free()


def test_kappa_precomputed_hessian_equals_stencil(golden_dir):
    """When the rays oversample the grid (nside 1024 on 1024^3: ~11 samples per cell) kappa_get_beam_properties evaluates
    the Hessian once per cell and the rays fetch it (tidal_field_kernel). Same float expressions: the maps must equal
    the per-sample 19-point stencil path bit for bit, and the oracle to fp32 rounding."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]))
    par = _par(t)
    par.grid_put(cb.GRID_NPOT, g["s1_npot"])
    par.update_halo()
    o.set_halo(g["s1_npot"])
    _, pos = cb.healpix.hp_shell_pixels(64, int(t["nside_base"]))          # 49152 rays x 16 samples >> 32768 cells
    rf = g["s6_kappa_rf"]
    a = cb.kappa_get_beam_properties(par, pos, rf)
    assert par.stage_ms("kappa_tidal")[1] == 0                              # (profiling off: no stage record)
    par.set_profiling(True)
    a2 = cb.kappa_get_beam_properties(par, pos, rf)
    assert par.stage_ms("kappa_tidal")[1] >= 1                              # the precompute pass really ran (per chunk)
    par.set_profiling(False)
    par.set_option("los_precompute", 0)
    b = cb.kappa_get_beam_properties(par, pos, rf)
    # same Hessians, same products; the chunks' partial sums are added in double and rounded once
    assert np.array_equal(a, a2)
    np.testing.assert_allclose(a, b, rtol=3e-7, atol=1e-7 * np.abs(b).max())
    assert (a == b).mean() > 0.99
    ref = o.kappa(g["s1_npot"], pos, rf)
    np.testing.assert_allclose(a, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
    par.free()


def test_imap_fast_painter_equals_exact(golden_dir):
    """The fp32-screened painter defers every sub-cell that is close to a pixel / shell edge to the double
    path, so hit counts must be IDENTICAL to the all-double kernel (imap.c:135-245) at any size."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    t = dict(t)
    n = 128
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    par = cb.ParamCoLoRe(t, n, seed=11)
    cb.create_cartesian_fields(par)
    cb.compute_physical_density_field(par)
    edges = np.linspace(0.12 * t["r_max"], 0.97 * t["r_max"], 13).astype(np.float32)
    res = {}
    for nside in (64, 256):
        par.set_imap(0, t["imap_tz_0"], t["imap_bz_0"], nside, edges[:-1], edges[1:])
        cb.compute_density_normalization(par)
        for exact in (1, 0):
            par.set_option("exact_math", exact)
            res[exact] = cb.imap_set_cartesian(par, 0)
        assert res[0][1].sum() > 5 * n ** 3 and np.array_equal(res[0][1], res[1][1]), nside
        np.testing.assert_allclose(res[0][0], res[1][0], rtol=2e-5, atol=1e-7 * np.abs(res[1][0]).max())
    par.free()


# ---------------------------------------------------------------------------------------- LPT
@pytest.mark.parametrize("name", ["ref_n32_1lpt_cic", "ref_n32_2lpt_tsc", "ref_n32_2lpt_ngp"])
def test_lpt_density_vs_reference(golden_dir, name):
    """lpt_1 / lpt_2 + deposit (density.c:37-188, 376-1031) from the reference's Gaussian field."""
    g, t = _load(golden_dir, name)
    n = int(t["n_grid"])
    o = Oracle(t, n)
    par = _par(t)
    par.set_option("lpt_interp_type", int(t["lpt_interp_type"]))
    par.set_option("keep_particles", 1)
    par.grid_put(cb.GRID_DENS, g["s1_dens_gauss"])
    cb.compute_physical_density_field(par)
    got = _real(par.grid_get(cb.GRID_DENS), n).astype(np.float64)
    ref = _real(g["s2_dens"], n).astype(np.float64)
    # particles: fp32 positions of O(1e3) Mpc/h built from fp32 FFTs -> a few ulps (1 ulp = 1.2e-4)
    d = g["s1_dens_gauss"].copy()
    pos_ref = o.lpt(d, int(t["dens_type"]), int(t["lpt_interp_type"]), want_pos=True)
    x, y, z = cb.lpt_get_particles(par)
    for a, b in zip((x, y, z), pos_ref):
        dd = np.abs(a.astype(np.float64) - b)
        dd = np.minimum(dd, np.abs(dd - par.l_box))          # periodic wrap
        assert dd.max() < 2e-3, dd.max()
    assert abs(got.sum()) < 0.5                               # mass conservation ("Total density", density.c:642)
    if int(t["lpt_interp_type"]) == 0:
        # NGP: a particle within rounding of a cell edge may land in the neighbour cell
        assert np.mean(got != ref) < 1e-3
    else:
        assert np.abs(got - ref).max() < 2e-4                 # CIC / TSC weights are continuous in position
    par.free()


def test_lpt_full_flow_runs(golden_dir):
    """2LPT at n=64 through the whole flow: density -> normalisation -> sources (BASELINE config 4 shape)."""
    g, t = _load(golden_dir, "ref_n32_2lpt_tsc")
    t = dict(t)
    n = 64
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    par = cb.ParamCoLoRe(t, n, dens_type=2, seed=5)
    par.set_option("lpt_interp_type", 1)
    cb.create_cartesian_fields(par)
    cb.compute_physical_density_field(par)
    dens = _real(par.grid_get(cb.GRID_DENS), n)
    assert dens.min() >= -1.0 and abs(float(dens.astype(np.float64).mean())) < 1e-5 and dens.std() > 0.01
    par.set_srcs(0, t["srcs_nz_0"] * 50, t["srcs_bz_0"])
    cb.compute_density_normalization(par)
    assert cb.srcs_set_cartesian(par)[0] > 1000
    par.free()


# ------------------------------------------------------------------- full-size property checks
def test_full_size_properties(golden_dir):
    """n_grid=512 (BASELINE config 2): properties that do not need the oracle at that size."""
    g, t = _load(golden_dir, "ref_n32_lognormal")
    n = 512
    t = dict(t)
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    par = cb.ParamCoLoRe(t, n, seed=7)
    mean, s2 = cb.create_cartesian_fields(par)
    assert abs(mean) < 1e-4 * np.sqrt(s2) and 0.05 < s2 < 5.0
    # Parseval-type check of sigma2 against an independent host reduction
    dens = par.grid_get(cb.GRID_DENS)[:, :, :n]
    assert abs(dens.astype(np.float64).var() / s2 - 1) < 1e-4
    # r2c(c2r(.)) round trip at full size
    x = par.grid_get(cb.GRID_NPOT)
    cb.fftw_wrap_r2c(par, cb.GRID_NPOT)
    cb.fftw_wrap_c2r(par, cb.GRID_NPOT)
    y = par.grid_get(cb.GRID_NPOT)
    assert np.abs(y[:, :, :n] / float(n) ** 3 - x[:, :, :n]).max() < 3e-5 * x[:, :, :n].std()
    par.grid_put(cb.GRID_NPOT, x)
    par.update_halo()
    cb.compute_physical_density_field(par)
    ln = par.grid_get(cb.GRID_DENS)[:, :, :n]
    assert ln.min() > -1.0 and abs(ln.astype(np.float64).mean()) < 0.05
    par.set_srcs(0, t["srcs_nz_0"] * 40, t["srcs_bz_0"])
    cb.compute_density_normalization(par)
    nsrc = cb.srcs_set_cartesian(par)[0]
    counts = cb.srcs_get_counts(par, 0)
    assert counts.sum() == nsrc and counts[:, :, n:].sum() == 0 and counts.min() >= 0
    pos, ipix = cb.srcs_get_cartesian(par, 0)
    assert ipix.min() >= 0 and ipix.max() < 48
    # sources are ordered by cell: the running cell index of consecutive sources never decreases
    dx = par.l_box / n
    cell = np.rint((pos[:, :3].astype(np.float64) + 0.5 * par.l_box) / dx).astype(np.int64)
    flat = cell[:, 0] + n * (cell[:, 1] + n * cell[:, 2])
    # (rint can move a source sitting exactly on a cell edge; allow a handful)
    assert np.mean(np.diff(flat) < 0) < 1e-3
    srcs = cb.srcs_get_local_properties(par, 0)
    assert srcs[:, 0].min() >= 0 and srcs[:, 0].max() <= 360.0 and np.abs(srcs[:, 1]).max() <= 90.0
    assert srcs[:, 2].max() < 0.5
    par.free()


# --------------------------------------------------------------------------------------- writer
def test_write_catalog_matches_io_c_formats(case, tmp_path):
    """clr_write_catalog (pinned chunked read-back + multi-threaded formatting) against the formats of write_catalog
    (io.c:1019-1236): the ASCII rows are exactly what C's "%d %E %E %E %E \\n" gives for the Src records (Python's %E is
    an independent correctly-rounded implementation), the FITS table holds the same floats bit for bit."""
    g, t, o, par = case
    _setup_sources(g, t, par)
    n = cb.srcs_set_cartesian(par)[0]
    srcs = cb.srcs_get_local_properties(par, 0)
    fa, ff = str(tmp_path / "cat.txt"), str(tmp_path / "cat.fits")
    for nt in (1, 5):
        cb.write_catalog(par, 0, fa, "ascii", n_threads=nt)
        lines = open(fa).read().split("\n")
        assert lines[0] == "#[1]type [2]RA, [3]dec, [4]z0, [5]dz_RSD " and lines[-1] == "" and len(lines) == n + 2
        want = ["0 %E %E %E %E " % tuple(float(v) for v in row[:4]) for row in srcs]
        assert lines[1:-1] == want
    cb.write_catalog(par, 0, ff, "fits")
    raw = open(ff, "rb").read()
    assert len(raw) % 2880 == 0
    cards = [raw[2880 + i:2880 + i + 80].decode() for i in range(0, 2880, 80)]
    keys = {c[:8].strip(): c[10:].split("/")[0].strip() for c in cards if c[8:10] == "= "}
    assert keys["XTENSION"] == "'BINTABLE'" and int(keys["NAXIS1"]) == 20 and int(keys["NAXIS2"]) == n and int(keys["TFIELDS"]) == 5
    assert [keys[f"TTYPE{i}"].strip("' ") for i in range(1, 6)] == ["TYPE", "RA", "DEC", "Z_COSMO", "DZ_RSD"]
    assert [keys[f"TFORM{i}"].strip("' ") for i in range(1, 6)] == ["1J", "1E", "1E", "1E", "1E"]
    rows = np.frombuffer(raw[5760:5760 + 20 * n], dtype=np.dtype([("t", ">i4"), ("f", ">f4", (4,))]))
    assert np.all(rows["t"] == 0) and np.array_equal(rows["f"].astype(np.float32), srcs[:, :4])
    assert raw[5760 + 20 * n:] == b"\0" * (len(raw) - 5760 - 20 * n)
