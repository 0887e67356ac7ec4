"""CPU suite: the C restatement (oracle/colore_oracle.c) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py -> oracle/_ref/ref_driver, OMP_NUM_THREADS=1).

With one thread the reference's MT19937 stream order is the plain (iz,iy,ix) loop order, which the
restatement reproduces, so every stage is compared BIT-EXACTLY (same compiler, same libm).
The only non-bit-exact comparison would be the intensity maps under OpenMP atomics; with one
thread they are exact too.
"""
import os

import numpy as np
import pytest

from oracle.oracle import NA, RNG_MT, Oracle, tables_from_dump

CASES = ["ref_n32_lognormal", "ref_n32_clip", "ref_n48_nosmooth", "ref_n32_1lpt_cic", "ref_n32_2lpt_tsc", "ref_n32_2lpt_ngp",
         "ref_n32_bias1", "ref_n32_bias3",       # the reference compiled with the other bias models (common.h:414-431)
         "ref_n32_nosmooth",                     # power-of-two grid without smoothing (runs on the GPU too)
         "ref_n32_lensing",                      # per-source lensing + density skewers + custom map (srcs.c:506-615, cstm.c)
         "ref_n32_fastlens",                     # -D_USE_FAST_LENSING build: lensing.c shells + srcs.c:666-721
         "ref_n32_gskw",                         # Gaussian skewers (beaming.c:55-66)
         "ref_n32_dense"]                        # ~120 sources per cell: gsl_ran_poisson's mu > 10 branch (common.c:187)


def _same(arr, g, key):
    """arr == g[key] bit for bit; large catalogues are stored as digests (make_golden.py:compact)."""
    if key in g:
        return np.array_equal(np.asarray(arr).ravel(), g[key].ravel())
    import hashlib
    a = np.ascontiguousarray(arr)
    return (a.size == int(g[key + "__size"][0]) and np.array_equal(a.ravel()[:4096], g[key + "__head"])
            and hashlib.sha256(a.tobytes()).digest() == g[key + "__sha256"].tobytes())


def _size(g, key):
    return g[key].size if key in g else int(g[key + "__size"][0])


def _bias_model(name):
    return 1 if "bias1" in name else (3 if "bias3" in name else 2)


@pytest.fixture(scope="module", params=CASES)
def case(request, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, request.param + ".npz")))
    t = tables_from_dump(g)
    o = Oracle(t, int(t["n_grid"]), nside_base=int(t["nside_base"]), bias_model=_bias_model(request.param))
    return g, t, o


def _fields(g, t, o):
    dk, pk = o.fill_modes(RNG_MT, int(t["seed"]))
    dens, npot = o.c2r(dk), o.c2r(pk)
    o.normalize_fields(dens, npot)
    return dens, npot


def test_gaussian_fields_bit_exact(case):
    g, t, o = case
    dens, npot = _fields(g, t, o)
    assert np.array_equal(dens, g["s1_dens_gauss"])
    assert np.array_equal(npot, g["s1_npot"])
    mean, s2 = o.sigma_dens(dens)
    assert s2 == g["s1_sigma2_gauss"][0]
    assert abs(mean) < 1e-6


def test_physical_density_and_normalisation(case):
    g, t, o = case
    dens = g["s1_dens_gauss"].copy()
    if int(t["dens_type"]) in (1, 2):       # lpt_1 / lpt_2 + deposit (density.c:376-1031)
        o.lpt(dens, int(t["dens_type"]), int(t["lpt_interp_type"]))
        n = o.n
        assert np.array_equal(dens[:, :, :n], g["s2_dens"][:, :, :n])
        assert abs(dens[:, :, :n].astype(np.float64).sum()) < 1e-2       # mass conservation of the deposit
        dens = g["s2_dens"].copy()
    else:
        o.lognormalize(dens, g["s1_sigma2_gauss"][0], clip=(int(t["dens_type"]) == 3))
        assert np.array_equal(dens, g["s2_dens"])
    npop = sum(1 for k in t if k.startswith("srcs_bz_"))
    bz = [t[f"srcs_bz_{i}"] for i in range(npop)]
    if "imap_bz_0" in t:
        bz.append(t["imap_bz_0"])
    if "cstm_bz_0" in t:                    # custom maps are normalised like one more population (density.c:1177-1178)
        bz.append(t["cstm_bz_0"])
    nm = o.density_normalization(dens, bz)
    for i in range(npop):
        assert np.array_equal(nm["norm"][i], g[f"s3_srcs_norm_{i}"])
        assert np.array_equal(nm["ends"][i], g[f"s3_srcs_norm_ends_{i}"])
    if "imap_bz_0" in t:
        assert np.array_equal(nm["norm"][npop], g["s3_imap_norm_0"])
    if "cstm_bz_0" in t:
        assert np.array_equal(nm["norm"][-1], g["s3_cstm_norm_0"])
        assert np.array_equal(nm["ends"][-1], g["s3_cstm_norm_ends_0"])
    assert np.array_equal(nm["zends"], g["s3_znorm_ends"])
    assert nm["hist_n"].sum() <= o.n ** 3


def test_sources_bit_exact(case):
    g, t, o = case
    dens, npot = g["s2_dens"], g["s1_npot"]
    o.set_halo(npot)
    npop = sum(1 for k in t if k.startswith("srcs_bz_"))
    for ipop in range(npop):
        ends = g[f"s3_srcs_norm_ends_{ipop}"]
        ns, tot = o.srcs_poisson(dens, t[f"srcs_nz_{ipop}"], t[f"srcs_bz_{ipop}"], g[f"s3_srcs_norm_{ipop}"],
                                 ends[0], ends[1], RNG_MT, int(t["seed"]), ipop)
        assert tot == _size(g, f"s4_srcs_ipix_{ipop}")
        assert ns[:, :, o.n:].sum() == 0          # padding columns never hold sources
        pos, ipix = o.srcs_place(npot, ns, RNG_MT, int(t["seed"]), ipop)
        assert _same(pos, g, f"s4_srcs_pos_{ipop}")
        assert _same(ipix, g, f"s4_srcs_ipix_{ipop}")
        srcs = o.srcs_local_properties(pos)
        if f"s5_srcs_cat_{ipop}" in g:
            ref = g[f"s5_srcs_cat_{ipop}"].reshape(-1, 9)
            assert np.array_equal(srcs[:, :6], ref[:, :6])
        else:
            # compact fixture: columns 6-8 (kappa, dra, ddec) are zero in a run without lensing
            full = np.zeros((srcs.shape[0], 9), np.float32)
            full[:, :6] = srcs[:, :6]
            assert _same(full, g, f"s5_srcs_cat_{ipop}")
        flags = g.get(f"s6_srcs_flags_{ipop}", np.zeros(3)).astype(int)
        if "s6_lens_r" in g:
            # fast-lensing build: RSD as usual, then shear / convergence / deflection interpolated from the shells
            o.srcs_beam_rsd(npot, pos, srcs)
            npp = g["s6_lens_npp"]
            nbeams = g["s6_lens_pos"].size // (3 * int(npp[-1]))
            data = np.concatenate([g[f"s6_lens_data_{i:03d}"] for i in range(len(npp))])
            srcs[:, 6:] = 0
            srcs, bad = o.srcs_fast_lensing(g["s6_lens_r"], g["s6_lens_nside"], npp, data, nbeams, pos, srcs)
            assert bad == 0
            assert np.array_equal(srcs, g[f"s6_srcs_cat_{ipop}"].reshape(-1, 9))
        elif f"s6_srcs_cat_{ipop}" in g and not flags.any():
            o.srcs_beam_rsd(npot, pos, srcs)
            assert np.array_equal(srcs[:, :6], g[f"s6_srcs_cat_{ipop}"].reshape(-1, 9)[:, :6])
        elif f"s6_srcs_cat_{ipop}" in g:
            # srcs.c:452-744 with per-source lensing and / or skewers: all 9 columns and both skewer arrays
            srcs[:, 6:] = 0
            srcs, dg, vs = o.srcs_beam_full(dens, npot, pos, srcs, g["s1_sigma2_gauss"][0], lensing=flags[0],
                                            skewers=flags[1], gaussian=flags[2])
            ref = g[f"s6_srcs_cat_{ipop}"].reshape(-1, 9)
            ncol = 9 if flags[0] else 6
            assert np.array_equal(srcs[:, :ncol], ref[:, :ncol])
            if flags[1]:
                assert np.array_equal(dg.ravel(), g[f"s6_srcs_dgskw_{ipop}"])
                assert np.array_equal(vs.ravel(), g[f"s6_srcs_vskw_{ipop}"])


def test_fast_lensing_shells_bit_exact(case):
    g, t, o = case
    if "s6_lens_r" not in g:
        pytest.skip("case has no lensing shells")
    o.set_halo(g["s1_npot"])
    npp = g["s6_lens_npp"]
    nr = len(npp)
    npix_hi = int(npp[-1])
    pos = g["s6_lens_pos"].reshape(-1, 3)
    nbeams = pos.shape[0] // npix_hi
    # the shell radii before the run: compute_lensing_spacing (cosmo.c:851-868), spacing in r
    r0 = ((np.arange(nr) + 1) * np.float32(np.float32(t["r_max"]) / np.float32(nr))).astype(np.float32)
    data, r_snap = o.lensing_shells(g["s1_npot"], r0, npp, pos, nbeams)
    assert np.array_equal(r_snap, g["s6_lens_r"])
    off = 0
    for i in range(nr):
        n5 = 5 * nbeams * int(npp[i])
        assert np.array_equal(data[off:off + n5], g[f"s6_lens_data_{i:03d}"]), i
        off += n5
    # pixel centres of the finest shell: NEST order inside each base pixel (common.c:493-502)
    lp, pp = o.shell_pixels(int(g["s6_lens_nside"][-1]))
    assert np.array_equal(pp, pos)


def test_custom_map_bit_exact(case):
    g, t, o = case
    if "s6_cstm_data_0" not in g:
        pytest.skip("case has no custom map")
    ends = g["s3_cstm_norm_ends_0"]
    pos = g["s6_cstm_pos_0"].reshape(-1, 3)
    data = o.cstm(g["s2_dens"], t["cstm_kz_0"], t["cstm_bz_0"], g["s3_cstm_norm_0"], ends[0], ends[1], pos)
    assert np.array_equal(data, g["s6_cstm_data_0"])


def test_maps_bit_exact(case):
    g, t, o = case
    if "s6_kappa_data" not in g:
        pytest.skip("case has no maps")
    dens, npot = g["s2_dens"], g["s1_npot"]
    o.set_halo(npot)
    r0, rf = g["s4_imap_r0_0"], g["s4_imap_rf_0"]
    nside = int(np.sqrt(g["s4_imap_data_0"].size / len(r0) / 12))
    ends = g["s3_imap_norm_ends_0"]
    data, nadd = o.imap(dens, npot, t["imap_tz_0"], t["imap_bz_0"], g["s3_imap_norm_0"], ends[0], ends[1],
                        nside, r0, rf)
    assert np.array_equal(nadd.ravel(), g["s4_imap_nadd_0"])
    assert np.array_equal(data.ravel(), g["s4_imap_data_0"])
    posk = g["s6_kappa_pos"].reshape(-1, 3)
    lp, pp = o.shell_pixels(int(np.sqrt(posk.shape[0] / 12)))
    assert np.array_equal(lp, g["s6_kappa_listpix"]) and np.array_equal(pp, posk)
    assert np.array_equal(o.kappa(npot, posk, g["s6_kappa_rf"]).ravel(), g["s6_kappa_data"])
    assert np.array_equal(o.isw(npot, posk, g["s6_isw_rf"]).ravel(), g["s6_isw_data"])


def test_table_lookup_edges(case):
    g, t, o = case
    # cosmo.c:30-38: clamps at r<=0 and beyond the table
    assert o.get_bg(-1.0, t["d1"], 1.0, t["d1"][-1]) == 1.0
    assert o.get_bg(1e9, t["z"], 0.0, t["z"][-1]) == t["z"][-1]
    r = 0.5 * (t["r"][10] + t["r"][11])
    assert abs(o.get_bg(r, t["z"], 0.0, 0.0) - 0.5 * (t["z"][10] + t["z"][11])) < 1e-12
    # cosmo.c:291-308: power-law extrapolation on both sides
    lo = o.pk_linear0(t["logkmin"] - 1.0)
    assert np.isclose(lo, t["pk_pk"][0] * 10 ** (-t["n_scal"]))
    hi = o.pk_linear0(t["logkmax"] + 1.0)
    assert np.isclose(hi, t["pk_pk"][-1] * 1e-3)
    assert len(t["z"]) == NA


def test_shim_poisson_moments():
    """gsl_ran_poisson is third-party arithmetic the repo has to restate (third_party/shim/gsl_shim.c):
    Knuth's product for mu <= 10, the gamma / binomial (BTPE) reduction above. No reference vectors exist for it,
    so its distribution is checked: mean and variance of 4000 draws at each mu within 5 sigma of mu."""
    import ctypes as C
    from oracle.oracle import RNG_PHILOX, build
    lib = C.CDLL(build())
    lib.orc_poisson_from_stream.restype = C.c_int
    lib.orc_poisson_from_stream.argtypes = [C.c_int, C.c_ulong, C.c_uint, C.c_ulonglong, C.c_double]
    n = 4000
    for mu in (0.03, 0.7, 9.5, 10.5, 37.0, 400.0, 12345.0):
        k = np.array([lib.orc_poisson_from_stream(RNG_PHILOX, 77, 3, i, mu) for i in range(n)], np.float64)
        assert k.min() >= 0
        assert abs(k.mean() - mu) < 5 * np.sqrt(mu / n), (mu, k.mean())
        # var of the sample variance of a Poisson: (mu + 2 mu^2 (n/(n-1))) / n  ~  (mu + 2 mu^2) / n
        assert abs(k.var(ddof=1) - mu) < 5 * np.sqrt((mu + 2 * mu * mu) / n), (mu, k.var(ddof=1))


@pytest.mark.parametrize("n", [16, 48])
def test_shim_fft_matches_numpy(golden_dir, n):
    """The reference's FFT is FFTW, which this image does not have: third_party/shim/fftw_shim.c restates the c2r / r2c
    definition (unnormalised, complex passes over z and y, half-complex x pass last). Checked against an
    independent implementation (numpy / pocketfft), including a NON-Hermitian spectrum like the one
    create_grids_fourier fills (fourier.c:325-345): the imaginary parts of the x-DC / x-Nyquist lines are dropped."""
    g = dict(np.load(os.path.join(golden_dir, "ref_n32_lognormal.npz")))
    t = tables_from_dump(g)
    o = Oracle(t, n)
    rng = np.random.default_rng(n)
    nc = n // 2 + 1
    ck = (rng.standard_normal((n, n, nc)) + 1j * rng.standard_normal((n, n, nc))).astype(np.complex64)
    got = o.c2r(ck.copy())[:, :, :n].astype(np.float64)
    ref = np.fft.irfftn(ck.astype(np.complex128), s=(n, n, n), axes=(0, 1, 2)) * float(n) ** 3
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()
    x = rng.standard_normal((n, n, n)).astype(np.float32)
    pad = np.zeros((n, n, 2 * nc), np.float32)
    pad[:, :, :n] = x
    gotk = o.r2c(pad).view(np.complex64).reshape(n, n, nc).astype(np.complex128)
    refk = np.fft.rfftn(x.astype(np.float64), axes=(0, 1, 2))
    assert np.abs(gotk - refk).max() < 2e-6 * np.abs(refk).max()


def test_shim_mt19937_known_answers():
    """gsl_rng_mt19937 restated in third_party/shim/gsl_shim.c: init_genrand seeding (seed 0 -> 4357), uniform = 32-bit
    output / 2^32. Known answers: the published MT19937 reference stream (seed 5489 -> 3499211612, 581869302,
    3890346734, ...; 10000th output 4123659995) and numpy's independent implementation for the run seed."""
    import ctypes as C
    from oracle.oracle import build
    lib = C.CDLL(build())
    lib.orc_mt19937_fill.argtypes = [C.c_ulong, C.c_long, C.c_void_p]

    def stream(seed, n):
        out = np.empty(n)
        lib.orc_mt19937_fill(seed, n, out.ctypes.data_as(C.c_void_p))
        return (out * 4294967296.0).astype(np.uint64)

    s = stream(5489, 10000)
    assert list(s[:3]) == [3499211612, 581869302, 3890346734] and s[9999] == 4123659995
    mt = np.random.MT19937()
    mt._legacy_seeding(1003)
    assert np.array_equal(stream(1003, 1000), mt.random_raw(1000))
    assert np.array_equal(stream(0, 5), stream(4357, 5))            # gsl_rng_set: seed 0 means 4357


def test_shim_philox_known_answers():
    """Philox4x32-10 (Salmon et al. 2011), the counter-based generator of the GPU path: the Random123 known-answer
    vectors. The CUDA implementation (clr_internal.cuh:clr_philox) is compared with this one in the GPU suite."""
    import ctypes as C
    from oracle.oracle import build
    lib = C.CDLL(build())
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c = (C.c_uint * 4)(*ctr)
        k = (C.c_uint * 2)(*key)
        out = (C.c_uint * 4)()
        lib.orc_philox(c, k, out)
        assert tuple(out) == want, [hex(v) for v in out]
