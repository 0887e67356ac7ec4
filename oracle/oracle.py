"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/liboracle.so (the CPU restatement).

Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package (colore_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NA = 5001
RNG_MT, RNG_PHILOX = 0, 1

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)


class OrcPar(C.Structure):
    _fields_ = [
        ("n_grid", C.c_int), ("nz_here", C.c_int), ("iz0_here", C.c_int), ("numk", C.c_int),
        ("do_smoothing", C.c_int), ("smooth_potential", C.c_int), ("bias_model_id", C.c_int),
        ("nside_base", C.c_int),
        ("l_box", C.c_float), ("pad_", C.c_float),
        ("pos_obs", C.c_double * 3),
        ("glob_idr", C.c_double), ("r2_smooth", C.c_double), ("prefac_lensing", C.c_double),
        ("fgrowth_0", C.c_double), ("hubble_0", C.c_double), ("OmegaM", C.c_double),
        ("n_scal", C.c_double),
        ("logkmin", C.c_double), ("logkmax", C.c_double), ("idlogk", C.c_double),
        ("r_max", C.c_double), ("sigma2_gauss", C.c_double),
        ("logkarr", c_double_p), ("pkarr", c_double_p),
        ("r_arr", c_double_p), ("z_arr", c_double_p), ("d1_arr", c_double_p), ("d2_arr", c_double_p),
        ("v1_arr", c_double_p), ("pd_arr", c_double_p), ("ih_arr", c_double_p),
        ("a_arr_a2r", c_double_p), ("r_arr_a2r", c_double_p),
        ("slice_left", c_float_p), ("slice_right", c_float_p),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
            os.path.join(HERE, "colore_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return so


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


class Oracle:
    """Holds the host tables (the inputs of the hot path) and calls the C restatement.

    ``tables``: dict with keys z, r, d1, d2, v1, pd, ih, a2r_a, a2r_r (NA doubles each),
    pk_logk, pk_pk, and scalars l_box, glob_idr, prefac_lensing, fgrowth_0, hubble_0, OmegaM,
    n_scal, r2_smooth, smooth_potential, do_smoothing, r_max, logkmin, logkmax, idlogk.
    """

    def __init__(self, tables: dict, n_grid: int, nz_here: int | None = None, iz0_here: int = 0,
                 bias_model: int = 2, nside_base: int = 2):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_pk_linear0.restype = C.c_double
        L.orc_pk_linear0.argtypes = [C.c_void_p, C.c_double]
        L.orc_get_bg.restype = C.c_double
        L.orc_get_bg.argtypes = [C.c_void_p, C.c_double, c_double_p, C.c_double, C.c_double]
        L.orc_bias_model.restype = C.c_double
        L.orc_bias_model.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_he_ang2pix.restype = C.c_long
        L.orc_he_ang2pix.argtypes = [C.c_long, C.c_double, C.c_double]
        L.orc_srcs_poisson.restype = C.c_long
        L.orc_hp_shell_pixels.restype = C.c_long
        L.orc_vec2pix_nest_via_ring.restype = C.c_long
        L.orc_vec2pix_nest_via_ring.argtypes = [C.c_long, c_double_p]
        L.orc_poisson_from_stream.argtypes = [C.c_int, C.c_ulong, C.c_uint, C.c_ulonglong, C.c_double]
        self.t = {k: (np.ascontiguousarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v)
                  for k, v in tables.items()}
        self.n = int(n_grid)
        self.nc = self.n // 2 + 1
        self.nz_here = self.n if nz_here is None else int(nz_here)
        p = OrcPar()
        p.n_grid, p.nz_here, p.iz0_here = self.n, self.nz_here, int(iz0_here)
        p.numk = len(self.t["pk_pk"])
        p.do_smoothing = int(self.t["do_smoothing"])
        p.smooth_potential = int(self.t["smooth_potential"])
        p.bias_model_id = bias_model
        p.nside_base = nside_base
        p.l_box = float(np.float32(self.t["l_box"]))
        for i in range(3):
            p.pos_obs[i] = float(self.t["pos_obs"])
        for k in ("glob_idr", "r2_smooth", "prefac_lensing", "fgrowth_0", "hubble_0", "OmegaM", "n_scal",
                  "logkmin", "logkmax", "idlogk", "r_max"):
            setattr(p, k, float(self.t[k]))
        p.sigma2_gauss = 0.0
        p.logkarr, p.pkarr = _dp(self.t["pk_logk"]), _dp(self.t["pk_pk"])
        p.r_arr, p.z_arr = _dp(self.t["r"]), _dp(self.t["z"])
        p.d1_arr, p.d2_arr, p.v1_arr = _dp(self.t["d1"]), _dp(self.t["d2"]), _dp(self.t["v1"])
        p.pd_arr, p.ih_arr = _dp(self.t["pd"]), _dp(self.t["ih"])
        p.a_arr_a2r, p.r_arr_a2r = _dp(self.t["a2r_a"]), _dp(self.t["a2r_r"])
        self.par = p
        self._halo = None

    # -- helpers ---------------------------------------------------------------------------
    @property
    def pp(self):
        return C.byref(self.par)

    def grid_shape(self):
        return (self.nz_here, self.n, 2 * self.nc)

    def set_halo(self, npot: np.ndarray, left: np.ndarray | None = None, right: np.ndarray | None = None):
        """z-halo planes of the potential (fourier.c:401-414). Single slab: periodic wrap."""
        if left is None:
            left = npot[self.nz_here - 1]
            right = npot[0]
        self._halo = (np.ascontiguousarray(left, np.float32), np.ascontiguousarray(right, np.float32))
        self.par.slice_left, self.par.slice_right = _fp(self._halo[0]), _fp(self._halo[1])

    # -- stages ----------------------------------------------------------------------------
    def pk_linear0(self, lgk: float) -> float:
        return self.lib.orc_pk_linear0(self.pp, lgk)

    def get_bg(self, r: float, tab: np.ndarray, f0: float, ff: float) -> float:
        return self.lib.orc_get_bg(self.pp, r, _dp(tab), f0, ff)

    def fill_modes(self, kind: int, seed: int):
        dens = np.zeros((self.nz_here, self.n, self.nc), np.complex64)
        npot = np.zeros_like(dens)
        self.lib.orc_fill_modes(self.pp, C.c_int(kind), C.c_ulong(seed), dens.ctypes.data_as(C.c_void_p),
                                npot.ctypes.data_as(C.c_void_p))
        return dens, npot

    def c2r(self, grid_c: np.ndarray) -> np.ndarray:
        """In-place c2r of a full n^3 grid; returns the padded real view [n][n][2*nc]."""
        assert grid_c.dtype == np.complex64 and grid_c.shape == (self.n, self.n, self.nc)
        self.lib.orc_c2r(C.c_int(self.n), grid_c.ctypes.data_as(C.c_void_p))
        return grid_c.view(np.float32).reshape(self.n, self.n, 2 * self.nc)

    def r2c(self, grid_r: np.ndarray) -> np.ndarray:
        assert grid_r.dtype == np.float32 and grid_r.shape == (self.n, self.n, 2 * self.nc)
        self.lib.orc_r2c(C.c_int(self.n), grid_r.ctypes.data_as(C.c_void_p))
        return grid_r.view(np.complex64).reshape(self.n, self.n, self.nc)

    def normalize_fields(self, dens: np.ndarray, npot: np.ndarray):
        self.lib.orc_normalize_fields(self.pp, _fp(dens), _fp(npot))

    def sigma_dens(self, dens: np.ndarray):
        out = np.zeros(2)
        self.lib.orc_sigma_dens(self.pp, _fp(dens), _dp(out))
        return out[0], out[1]

    def lognormalize(self, dens: np.ndarray, sigma2: float, clip: bool = False):
        self.par.sigma2_gauss = float(sigma2)
        self.lib.orc_lognormalize(self.pp, _fp(dens), C.c_int(int(clip)))

    def lpt(self, dens: np.ndarray, order: int, interp_type: int, want_pos: bool = False):
        """lpt_1 / lpt_2 (density.c:376-1031) in place on the Gaussian field; optionally the particles."""
        pos = np.zeros((3, self.nz_here * self.n * self.n), np.float32) if want_pos else None
        self.lib.orc_lpt(self.pp, C.c_int(order), C.c_int(interp_type), _fp(dens),
                         _fp(pos) if want_pos else None)
        return pos

    def density_normalization(self, dens: np.ndarray, bz_tabs: list):
        npop = len(bz_tabs)
        nz = self.lib.orc_norm_nz(self.pp)
        tabs = [np.ascontiguousarray(b, np.float64) for b in bz_tabs]
        arr = (c_double_p * max(npop, 1))(*[_dp(b) for b in tabs])
        norm = np.zeros((npop, NA))
        ends = np.zeros(2 * max(npop, 1))
        zends = np.zeros(2)
        hn = np.zeros(nz, np.uint64)
        hz = np.zeros(nz)
        hb = np.zeros((max(npop, 1), nz))
        self.lib.orc_density_normalization(self.pp, _fp(dens), C.c_int(npop), arr, _dp(norm), _dp(ends),
                                           _dp(zends), hn.ctypes.data_as(C.c_void_p), _dp(hz), _dp(hb))
        return dict(norm=norm, ends=ends.reshape(-1, 2)[:npop], zends=zends, hist_n=hn, hist_z=hz,
                    hist_b=hb[:npop])

    def srcs_poisson(self, dens, nz_tab, bz_tab, norm_tab, norm_0, norm_f, kind, seed, ipop=0):
        ns = np.zeros(self.grid_shape(), np.int32)
        tot = self.lib.orc_srcs_poisson(self.pp, _fp(dens), _dp(nz_tab), _dp(bz_tab), _dp(norm_tab),
                                        C.c_double(norm_0), C.c_double(norm_f), C.c_int(kind),
                                        C.c_ulong(seed), C.c_int(ipop), ns.ctypes.data_as(C.c_void_p))
        return ns, int(tot)

    def srcs_place(self, npot, nsources, kind, seed, ipop=0):
        n = int(nsources.sum())
        pos = np.zeros((n, 4), np.float32)
        ipix = np.zeros(n, np.int32)
        self.lib.orc_srcs_place(self.pp, _fp(npot), nsources.ctypes.data_as(C.c_void_p), C.c_int(kind),
                                C.c_ulong(seed), C.c_int(ipop), _fp(pos), ipix.ctypes.data_as(C.c_void_p))
        return pos, ipix

    def srcs_local_properties(self, pos):
        n = pos.shape[0]
        srcs = np.zeros((n, 9), np.float32)
        self.lib.orc_srcs_local_properties(self.pp, _fp(pos), C.c_long(n), _fp(srcs))
        return srcs

    def srcs_beam_rsd(self, npot, pos, srcs, pre=True, post=True):
        self.lib.orc_srcs_beam_rsd(self.pp, _fp(npot), _fp(pos), C.c_long(pos.shape[0]), _fp(srcs),
                                   C.c_int(int(pre)), C.c_int(int(post)))
        return srcs

    def srcs_beam_full(self, dens, npot, pos, srcs, sigma2, lensing=False, skewers=False, gaussian=False,
                       dg_skw=None, v_skw=None, pre=True, post=True):
        """srcs.c:425-744 (RSD + skewers + per-source lensing). Returns (srcs, dg_skw, v_skw)."""
        n = pos.shape[0]
        nr = self.n // 2
        self.par.sigma2_gauss = float(sigma2)
        if skewers and dg_skw is None:
            dg_skw = np.zeros((n, nr), np.float32)
            v_skw = np.zeros((n, nr), np.float32)
        self.lib.orc_srcs_get_beam_properties(self.pp, _fp(dens), _fp(npot), _fp(pos), C.c_long(n), _fp(srcs),
                                              C.c_int(int(lensing)), C.c_int(int(skewers)), C.c_int(int(gaussian)),
                                              _fp(dg_skw) if skewers else None, _fp(v_skw) if skewers else None,
                                              C.c_int(int(pre)), C.c_int(int(post)))
        return srcs, dg_skw, v_skw

    def lensing_shells(self, npot, r_sh, npp, pos, nbeams, data=None, snap=True):
        """lensing.c:76-250. Returns (data flat, r_sh snapped). data: shells concatenated, [nbeams][5*npp[ir]] each."""
        r_sh = np.ascontiguousarray(r_sh, np.float32).copy()
        npp = np.ascontiguousarray(npp, np.int64)
        if data is None:
            data = np.zeros(5 * nbeams * int(npp.sum()), np.float32)
        self.lib.orc_lensing_get_beam_properties(self.pp, _fp(npot), C.c_int(nbeams), C.c_int(len(r_sh)), _fp(r_sh),
                                                 npp.ctypes.data_as(C.c_void_p), _dp(pos), _fp(data), C.c_int(int(snap)))
        return data, r_sh

    def srcs_fast_lensing(self, r_sh, nside_sh, npp, data, nbeams, pos, srcs, node=0, nnodes=1):
        """srcs.c:666-723: interpolation of the shell quantities onto the sources (updates srcs columns 4-8)."""
        r_sh = np.ascontiguousarray(r_sh, np.float32)
        nside_sh = np.ascontiguousarray(nside_sh, np.int64)
        npp = np.ascontiguousarray(npp, np.int64)
        self.lib.orc_srcs_fast_lensing.restype = C.c_long
        bad = self.lib.orc_srcs_fast_lensing(self.pp, C.c_int(len(r_sh)), _fp(r_sh), nside_sh.ctypes.data_as(C.c_void_p),
                                             npp.ctypes.data_as(C.c_void_p), _fp(data), C.c_int(nbeams), C.c_int(node),
                                             C.c_int(nnodes), _fp(pos), C.c_long(pos.shape[0]), _fp(srcs))
        return srcs, int(bad)

    def cstm(self, dens, kz_tab, bz_tab, norm_tab, norm_0, norm_f, pos, data=None):
        """cstm.c:68-145: custom projected map for the unit vectors pos[npix][3]."""
        npix = pos.shape[0]
        if data is None:
            data = np.zeros(npix, np.float32)
        self.lib.orc_cstm_get_beam_properties(self.pp, _fp(dens), _dp(kz_tab), _dp(bz_tab), _dp(norm_tab),
                                              C.c_double(norm_0), C.c_double(norm_f), C.c_long(npix), _dp(pos), _fp(data))
        return data

    def imap(self, dens, npot, tz_tab, bz_tab, norm_tab, norm_0, norm_f, nside, r0, rf):
        nr = len(r0)
        npix = 12 * nside * nside
        data = np.zeros((nr, npix), np.float32)
        nadd = np.zeros((nr, npix), np.int32)
        r0 = np.ascontiguousarray(r0, np.float32)
        rf = np.ascontiguousarray(rf, np.float32)
        self.lib.orc_imap_set_cartesian(self.pp, _fp(dens), _fp(npot), _dp(tz_tab), _dp(bz_tab), _dp(norm_tab),
                                        C.c_double(norm_0), C.c_double(norm_f), C.c_int(nside), C.c_int(nr),
                                        _fp(r0), _fp(rf), _fp(data), nadd.ctypes.data_as(C.c_void_p))
        return data, nadd

    def shell_pixels(self, nside, nside_base=None, node=0, nnodes=1):
        nb = self.par.nside_base if nside_base is None else nside_base
        n = self.lib.orc_hp_shell_pixels(C.c_int(nside), C.c_int(nb), C.c_int(node), C.c_int(nnodes), None, None)
        lp = np.zeros(n, np.int64)
        pos = np.zeros((n, 3))
        self.lib.orc_hp_shell_pixels(C.c_int(nside), C.c_int(nb), C.c_int(node), C.c_int(nnodes),
                                     lp.ctypes.data_as(C.c_void_p), _dp(pos))
        return lp, pos

    def kappa(self, npot, pos, rf, data=None):
        rf = np.ascontiguousarray(rf, np.float32)
        npix = pos.shape[0]
        if data is None:
            data = np.zeros((len(rf), npix), np.float32)
        self.lib.orc_kappa_get_beam_properties(self.pp, _fp(npot), C.c_long(npix), _dp(pos), C.c_int(len(rf)),
                                               _fp(rf), _fp(data))
        return data

    def isw(self, npot, pos, rf, data=None):
        rf = np.ascontiguousarray(rf, np.float32)
        npix = pos.shape[0]
        if data is None:
            data = np.zeros((len(rf), npix), np.float32)
        self.lib.orc_isw_get_beam_properties(self.pp, _fp(npot), C.c_long(npix), _dp(pos), C.c_int(len(rf)),
                                             _fp(rf), _fp(data))
        return data

    def he_ang2pix(self, nside, cth, phi):
        return self.lib.orc_he_ang2pix(nside, cth, phi)


def tables_from_dump(d) -> dict:
    """Host tables written by oracle/ref_driver.c (the reference's own cosmo.c output).

    ``d`` is a dump directory of .npy files or a mapping name -> array (a loaded golden .npz)."""
    if not isinstance(d, str):
        return tables_from_arrays(d)
    return tables_from_arrays({f[:-4]: np.load(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".npy")})


def tables_from_arrays(arrs) -> dict:
    ld = lambda n: np.asarray(arrs[n])  # noqa: E731
    sc = ld("scalars")
    t = dict(z=ld("tab_z"), r=ld("tab_r"), d1=ld("tab_d1"), d2=ld("tab_d2"), v1=ld("tab_v1"), pd=ld("tab_pd"),
             ih=ld("tab_ih"), a2r_a=ld("tab_a2r_a"), a2r_r=ld("tab_a2r_r"), pk_logk=ld("pk_logk"),
             pk_pk=ld("pk_pk"))
    names = ["l_box", "pos_obs", "glob_idr", "prefac_lensing", "fgrowth_0", "hubble_0", "OmegaM", "n_scal",
             "r2_smooth", "smooth_potential", "do_smoothing", "r_max", "logkmin", "logkmax", "idlogk", "numk",
             "n_grid", "seed", "nside_base", "dens_type", "sigma2_analytic", "r_min", "z_min", "z_max",
             "lpt_interp_type", "lpt_buffer_fraction"]
    for i, k in enumerate(names):
        t[k] = float(sc[i])
    for k in list(arrs.keys()):
        if k.startswith("tab_srcs_") or k.startswith("tab_imap_") or k.startswith("tab_cstm_"):
            t[k[4:]] = np.asarray(arrs[k])
    return t
