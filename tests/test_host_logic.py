"""CPU suite: the host-side mirror of the reference's set-up code, against the tables the UNMODIFIED reference
wrote into the golden fixtures (tests/golden/make_golden.py), and the C ABI surface.

* colore_b200.cosmo.cosmo_set  vs  cosmo_set / pk_linear_set (cosmo.c:424-700) tables tab_* of the fixtures
* colore_b200.healpix.hp_shell_pixels  vs  hp_shell_alloc (common.c:505-552) pixel lists / unit vectors
* colore_b200.cosmo.choose_nside_base  vs  io.c:224-244
* every function declared in include/colore_b200.h is exported by libcolore_b200.so (no compute call)
"""
import os
import shutil
import tempfile

import numpy as np
import pytest

import colore_b200 as cb
from colore_b200.inputs import RunConfig, write_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tables_like_golden(cfg):
    d = tempfile.mkdtemp(prefix="clr_host_")
    try:
        paths = write_inputs(d, cfg)
        k, pk = np.loadtxt(paths["pk"], unpack=True)
        z, nz = np.loadtxt(paths["nz0"], unpack=True)
        _, bz = np.loadtxt(paths["bz0"], unpack=True)
        tz = bz_im = None
        if cfg.imap_nside > 0:
            zt, tz = np.loadtxt(paths["tz"], unpack=True)
            _, bz_im = np.loadtxt(paths["bz_im"], unpack=True)
    finally:
        shutil.rmtree(d)
    kw = {}
    if tz is not None:
        kw = dict(tz_tabs=[(zt, tz)], bz_imap_tabs=[(zt, bz_im)])
    return cb.cosmo.cosmo_set(cfg, k, pk, [(z, nz)], [(z, bz)], **kw)


def test_cosmo_tables_match_reference():
    """Background tables on the NA r-grid and the normalised P(k) as the reference's cosmo_set builds them."""
    g = np.load(os.path.join(GOLD, "ref_n32_lognormal.npz"))
    cfg = RunConfig(n_grid=32, dens_type=0, nz_amplitude=60.0, imap_nside=8, imap_nchannels=4, kappa_nside=8,
                    isw_nside=8, seed=1003)                       # the fixture's configuration (make_golden.py)
    t = _tables_like_golden(cfg)
    # measured agreement: ~1e-8 of the table's largest entry (two quadrature implementations of the same integrals); pd =
    # D H (f - 1) cancels at high redshift, where f -> 1: 3e-5
    for key, tab, rtol in (("r", "tab_r", 1e-7), ("z", "tab_z", 1e-7), ("d1", "tab_d1", 2e-7), ("d2", "tab_d2", 2e-7),
                           ("v1", "tab_v1", 2e-7), ("pd", "tab_pd", 1e-4), ("ih", "tab_ih", 1e-7),
                           ("a2r_r", "tab_a2r_r", 1e-7)):
        ref = g[tab]
        np.testing.assert_allclose(t[key], ref, rtol=rtol, atol=rtol * np.abs(ref).max(), err_msg=key)
    # P(k) table normalised to sigma_8 (pk_linear_set, cosmo.c:310-422)
    np.testing.assert_allclose(t["pk_logk"], g["pk_logk"], rtol=1e-12)
    np.testing.assert_allclose(t["pk_pk"], g["pk_pk"], rtol=1e-6)       # sigma_8 integral: 4e-8 between the two quadratures
    # population tables: dN/dz dOmega -> n(r), b(r), T(r) (cosmo.c:549-629)
    for key, tab in (("srcs_nz_0", "tab_srcs_nz_0"), ("srcs_bz_0", "tab_srcs_bz_0"), ("imap_tz_0", "tab_imap_tz_0"),
                     ("imap_bz_0", "tab_imap_bz_0")):
        ref = g[tab]
        np.testing.assert_allclose(np.nan_to_num(t[key]), np.nan_to_num(ref), rtol=1e-7, atol=1e-7 * np.abs(np.nan_to_num(ref)).max(),
                                   err_msg=key)


def test_hp_shell_pixels_match_reference():
    """NEST ids and pix2vec_nest unit vectors of the pixels a rank owns (common.c:517-540), bit for bit."""
    g = np.load(os.path.join(GOLD, "ref_n32_lognormal.npz"))
    nside = int(np.sqrt(g["s6_kappa_listpix"].size / 12))
    listpix, pos = cb.healpix.hp_shell_pixels(nside, 2)
    assert np.array_equal(listpix, g["s6_kappa_listpix"])
    assert np.array_equal(pos.ravel(), g["s6_kappa_pos"])
    # ownership by base pixel for several ranks: a partition of the sphere
    alls = np.concatenate([cb.healpix.hp_shell_pixels(16, 2, node=r, nnodes=3)[0] for r in range(3)])
    assert np.array_equal(np.sort(alls), np.arange(12 * 16 * 16))
    v = cb.healpix.pix2vec_nest(16, np.arange(12 * 16 * 16))
    np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-14)
    assert abs(v.sum(axis=0)).max() < 1e-10                         # equal-area pixels, symmetric


def test_choose_nside_base():
    """io.c:224-244: the coarsest nside (>= 2) whose 12 nside^2 base pixels split over the ranks with < 20 % load
    imbalance (or exactly). Expected values worked out by hand from that rule."""
    for nnodes, want in ((1, 2), (2, 2), (5, 2), (7, 2), (10, 4), (48, 2), (49, 8), (96, 4)):
        assert cb.cosmo.choose_nside_base(nnodes) == want, nnodes


def test_c_abi_exports_every_declared_symbol():
    """include/colore_b200.h is the drop-in boundary: the shared library must export all of it."""
    lib = cb.load()
    names = cb.declared_symbols()
    assert len(names) >= 35 and "clr_create_cartesian_fields" in names and "clr_comm_p2p" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.clr_version() >= 100
    # no GPU here: creating a context must fail loudly instead of falling back to the CPU
    if lib.clr_device_count() == 0:
        with pytest.raises(cb.ColoreError):
            cb.ParamCoLoRe({}, 32)
