"""Host-side mirror of the reference's run flow for the GPU path.

Function names, order and meaning follow main.c:50-147 / common.h:436-543: ``read_run_params``
-> ``create_cartesian_fields`` -> ``compute_physical_density_field`` ->
``compute_density_normalization`` -> ``srcs_set_cartesian`` / ``imap_set_cartesian`` ->
``srcs_distribute`` -> ``srcs_get_local_properties`` -> ``get_beam_properties``. Each one is a thin
call into the C ABI (include/colore_b200.h); errors raise ``ColoreError`` (the reference's
``report_error(1, ...)`` + exit).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NA, ClrParams, ColoreError, check

GRID_DENS, GRID_NPOT = 0, 1


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _tab(t, key):
    a = np.ascontiguousarray(t[key], dtype=np.float64)
    assert a.shape == (NA,), key
    return a


class ParamCoLoRe:
    """The state object of a run (common.h:220-381), GPU-resident.

    ``tables``: dict as produced by colore_b200.cosmo.cosmo_set (or the reference's own tables).
    """

    def __init__(self, tables: dict, n_grid: int, *, dens_type: int = 0, bias_model: int = 2, seed: int = 1003,
                 nside_base: int = 2, nz_here: int | None = None, iz0_here: int = 0, device: int = 0):
        self.lib = _lib.load()
        if self.lib.clr_device_count() <= device:
            raise ColoreError("no CUDA device visible: colore_b200 has no CPU fallback")
        self.t = tables
        self.n_grid = int(n_grid)
        self.nc = self.n_grid // 2 + 1
        self.nz_here = self.n_grid if nz_here is None else int(nz_here)
        self.iz0_here = int(iz0_here)
        self.seed = int(seed)
        self.dens_type = int(dens_type)
        self._keep = {}
        p = ClrParams()
        p.n_grid, p.nz_here, p.iz0_here = self.n_grid, self.nz_here, self.iz0_here
        p.dens_type, p.bias_model = self.dens_type, int(bias_model)
        p.do_smoothing, p.smooth_potential = int(tables["do_smoothing"]), int(tables["smooth_potential"])
        p.nside_base = int(nside_base)
        p.seed_rng = self.seed
        p.l_box = float(np.float32(tables["l_box"]))
        for i in range(3):
            p.pos_obs[i] = float(tables["pos_obs"])
        for k in ("r2_smooth", "prefac_lensing", "fgrowth_0", "hubble_0", "OmegaM", "n_scal", "r_max", "glob_idr",
                  "logkmin", "logkmax", "idlogk"):
            setattr(p, k, float(tables[k]))
        pk_logk = np.ascontiguousarray(tables["pk_logk"], np.float64)
        pk_pk = np.ascontiguousarray(tables["pk_pk"], np.float64)
        p.numk = len(pk_pk)
        self._keep["pk"] = (pk_logk, pk_pk)
        dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731
        p.logkarr, p.pkarr = dp(pk_logk), dp(pk_pk)
        for field, key in (("r_arr_r2z", "r"), ("z_arr_r2z", "z"), ("growth_d_arr", "d1"), ("growth_d2_arr", "d2"),
                           ("growth_v_arr", "v1"), ("growth_pd_arr", "pd"), ("ihub_arr", "ih"),
                           ("a_arr_a2r", "a2r_a"), ("r_arr_a2r", "a2r_r")):
            a = _tab(tables, key)
            self._keep[field] = a
            setattr(p, field, dp(a))
        self.l_box = p.l_box
        self.r_max = float(tables["r_max"])
        self.ctx = C.c_void_p()
        check(self.lib.clr_create(C.byref(p), C.c_int(device), C.byref(self.ctx)))
        self.sigma2_gauss = 0.0
        self.mean_gauss = 0.0
        self.n_srcs = 0
        self.n_imap = 0
        self.nsources = {}
        self.imap_shells = {}

    # -- lifetime ---------------------------------------------------------------------------
    def free(self):
        """param_colore_free (io.c:1247-1348)."""
        if self.ctx:
            self.lib.clr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def synchronize(self):
        check(self.lib.clr_synchronize(self.ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.clr_launch_count(self.ctx))

    def grid_shape(self):
        """Host-side (reference) layout of a grid: rows of 2*(n/2+1) floats (fourier.c:46-51)."""
        return (self.nz_here, self.n_grid, 2 * self.nc)

    def grid_pitch(self) -> int:
        """Floats per row of the DEVICE grids (rows are padded to 64 bytes; see clr_grid_pitch)."""
        v = C.c_longlong()
        check(self.lib.clr_grid_pitch(self.ctx, C.byref(v)))
        return int(v.value)

    def grid_device_ptr(self, which: int) -> int:
        p = C.c_void_p()
        check(self.lib.clr_grid_device_ptr(self.ctx, C.c_int(which), C.byref(p)))
        return int(p.value)

    # -- populations ------------------------------------------------------------------------
    def set_srcs(self, ipop: int, nz_arr, bz_arr):
        a, b = np.ascontiguousarray(nz_arr, np.float64), np.ascontiguousarray(bz_arr, np.float64)
        check(self.lib.clr_set_srcs(self.ctx, C.c_int(ipop), _vp(a), _vp(b)))
        self.n_srcs = max(self.n_srcs, ipop + 1)

    def set_imap(self, ipop: int, tz_arr, bz_arr, nside: int, r0, rf):
        a, b = np.ascontiguousarray(tz_arr, np.float64), np.ascontiguousarray(bz_arr, np.float64)
        r0 = np.ascontiguousarray(r0, np.float32)
        rf = np.ascontiguousarray(rf, np.float32)
        check(self.lib.clr_set_imap(self.ctx, C.c_int(ipop), _vp(a), _vp(b), C.c_int(nside), C.c_int(len(r0)),
                                    _vp(r0), _vp(rf)))
        self.n_imap = max(self.n_imap, ipop + 1)
        self.imap_shells[ipop] = (nside, len(r0))

    def set_cstm(self, ipop: int, kz_arr, bz_arr):
        """Custom projected tracer (cstm.c): K(z) and b(z) tables (cosmo.c:631-717)."""
        a, b = np.ascontiguousarray(kz_arr, np.float64), np.ascontiguousarray(bz_arr, np.float64)
        check(self.lib.clr_set_cstm(self.ctx, C.c_int(ipop), _vp(a), _vp(b)))

    # -- grids ------------------------------------------------------------------------------
    def grid_put(self, which: int, host: np.ndarray):
        """Upload a grid in the reference layout: float32 [nz][n][2*nc] or complex64 [nz][n][nc]."""
        if host.dtype == np.complex64:
            host = host.view(np.float32)
        host = np.ascontiguousarray(host, np.float32)
        assert host.size == np.prod(self.grid_shape())
        check(self.lib.clr_grid_put(self.ctx, C.c_int(which), _vp(host)))

    def grid_get(self, which: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.grid_shape(), np.float32)
        check(self.lib.clr_grid_get(self.ctx, C.c_int(which), _vp(out)))
        return out

    def set_option(self, name: str, value: int):
        check(self.lib.clr_set_option(self.ctx, name.encode(), C.c_int(int(value))))

    def set_sigma2_gauss(self, s2: float):
        self.sigma2_gauss = float(s2)
        check(self.lib.clr_set_sigma2_gauss(self.ctx, C.c_double(s2)))

    def update_halo(self):
        check(self.lib.clr_update_halo(self.ctx))

    # -- profiling --------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        check(self.lib.clr_set_profiling(self.ctx, C.c_int(int(on))))

    def stage_ms(self, stage: str):
        ms, nl = C.c_float(), C.c_int()
        check(self.lib.clr_get_stage_ms(self.ctx, stage.encode(), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def timer_start(self):
        check(self.lib.clr_timer_start(self.ctx))

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.clr_timer_stop_ms(self.ctx, C.byref(ms)))
        return ms.value


# ---- the reference's run-flow functions (common.h:436-543) ---------------------------------------

def fill_modes(par: ParamCoLoRe, seed: int | None = None):
    """create_grids_fourier (fourier.c:285-359)."""
    check(par.lib.clr_fill_modes(par.ctx, C.c_uint32(par.seed if seed is None else seed)))


def fftw_wrap_c2r(par: ParamCoLoRe, which: int):
    """fourier.c:81-102, in place on a device grid."""
    check(par.lib.clr_fft_c2r(par.ctx, C.c_int(which)))


def fftw_wrap_r2c(par: ParamCoLoRe, which: int):
    """fourier.c:104-125."""
    check(par.lib.clr_fft_r2c(par.ctx, C.c_int(which)))


def create_cartesian_fields(par: ParamCoLoRe, inject: bool = False):
    """fourier.c:361-423. ``inject``: Fourier modes were uploaded with grid_put (reference white noise)."""
    out = (C.c_double * 2)()
    check(par.lib.clr_create_cartesian_fields(par.ctx, C.c_uint32(par.seed), C.c_int(int(inject)), out))
    par.mean_gauss, par.sigma2_gauss = out[0], out[1]
    return out[0], out[1]


def normalize_fields(par: ParamCoLoRe):
    """fourier.c:381-416 on grids that are already in real space."""
    out = (C.c_double * 2)()
    check(par.lib.clr_normalize_fields(par.ctx, out))
    par.mean_gauss, par.sigma2_gauss = out[0], out[1]
    return out[0], out[1]


def compute_physical_density_field(par: ParamCoLoRe):
    """density.c:1105-1126."""
    check(par.lib.clr_compute_physical_density_field(par.ctx))


def lpt_get_particles(par: ParamCoLoRe):
    """Particles of the last LPT density (x, y, z arrays); needs par.set_option("keep_particles", 1)."""
    n = par.nz_here * par.n_grid * par.n_grid
    x, y, z = (np.empty(n, np.float32) for _ in range(3))
    check(par.lib.clr_lpt_get_particles(par.ctx, _vp(x), _vp(y), _vp(z)))
    return x, y, z


def lpt_exchange_counts(par: ParamCoLoRe):
    """(sent, received): particles this rank shipped to / got from other slabs in the last LPT density
    (share_particles, density.c:191-374)."""
    a, b = C.c_longlong(), C.c_longlong()
    check(par.lib.clr_lpt_exchange_counts(par.ctx, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def compute_density_normalization(par: ParamCoLoRe):
    """density.c:1227-1393."""
    check(par.lib.clr_compute_density_normalization(par.ctx))


def get_norm(par: ParamCoLoRe, kind: int, ipop: int):
    norm = np.empty(NA)
    ends, zends = np.empty(2), np.empty(2)
    check(par.lib.clr_get_norm(par.ctx, C.c_int(kind), C.c_int(ipop), _vp(norm), _vp(ends), _vp(zends)))
    return norm, ends, zends


def set_norm(par: ParamCoLoRe, kind: int, ipop: int, norm, ends):
    norm = np.ascontiguousarray(norm, np.float64)
    ends = np.ascontiguousarray(ends, np.float64)
    check(par.lib.clr_set_norm(par.ctx, C.c_int(kind), C.c_int(ipop), _vp(norm), _vp(ends)))


def srcs_set_cartesian(par: ParamCoLoRe, seed: int | None = None):
    """srcs.c:285-294 (all populations); also fills the Src records (srcs.c:386-416)."""
    for ipop in range(par.n_srcs):
        n = C.c_longlong()
        check(par.lib.clr_srcs_set_cartesian(par.ctx, C.c_int(ipop), C.c_uint32(par.seed if seed is None else seed),
                                             C.byref(n)))
        par.nsources[ipop] = int(n.value)
    return dict(par.nsources)


def srcs_distribute(par: ParamCoLoRe, by_pixel: bool = False, beam_first: bool = False):
    """srcs.c:296-384. Default: every GPU keeps (and writes) the sources of its own z slab. ``by_pixel``: the
    reference's routing, source -> rank ipix % NNodes with the order preserved (clr_srcs_distribute)."""
    if not by_pixel:
        return dict(par.nsources)
    for ipop in range(par.n_srcs):
        n = C.c_longlong()
        check(par.lib.clr_srcs_distribute(par.ctx, C.c_int(ipop), C.c_int(int(beam_first)), C.byref(n)))
        par.nsources[ipop] = int(n.value)
    return dict(par.nsources)


def srcs_get_counts(par: ParamCoLoRe, ipop: int = 0) -> np.ndarray:
    out = np.empty(par.grid_shape(), np.int32)
    check(par.lib.clr_srcs_get_counts(par.ctx, C.c_int(ipop), _vp(out)))
    return out


def srcs_get_cartesian(par: ParamCoLoRe, ipop: int = 0):
    """CatalogCartesian (common.h:163-167) of population ipop -> (pos[n,4], ipix[n])."""
    n = par.nsources.get(ipop, 0)
    pos = np.empty((n, 4), np.float32)
    ipix = np.empty(n, np.int32)
    check(par.lib.clr_srcs_get_cartesian(par.ctx, C.c_int(ipop), _vp(pos), _vp(ipix)))
    return pos, ipix


def srcs_get_local_properties(par: ParamCoLoRe, ipop: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """Catalog.srcs (common.h:169-194): float32 [n,9] = ra, dec, z0, dz_rsd, e1, e2, kappa, dra, ddec."""
    n = par.nsources.get(ipop, 0)
    if out is None:
        out = np.empty((n, 9), np.float32)
    check(par.lib.clr_srcs_get_local_properties(par.ctx, C.c_int(ipop), _vp(out)))
    return out


def srcs_beams(par: ParamCoLoRe):
    """srcs_beams_preproc / get_beam_properties (RSD part) / postproc (srcs.c:425-443,486-504,656-662)."""
    for ipop in range(par.n_srcs):
        check(par.lib.clr_srcs_beam_rsd(par.ctx, C.c_int(ipop)))


def srcs_get_beam_properties(par: ParamCoLoRe, ipop: int = 0, *, lensing: bool = False, skewers: bool = False,
                             gaussian_skewers: bool = False, rsd_done: bool = False):
    """srcs_beams_preproc + srcs_get_beam_properties + srcs_beams_postproc (srcs.c:425-744): RSD under beaming,
    per-source lensing (e1, e2, kappa, dra, ddec) and skewers. Returns (srcs[n,9], dg_skw[n,nr] | None, v_skw | None)."""
    check(par.lib.clr_srcs_get_beam_properties(par.ctx, C.c_int(ipop), C.c_int(int(lensing)), C.c_int(int(skewers)),
                                               C.c_int(int(gaussian_skewers)), C.c_int(int(rsd_done))))
    srcs = srcs_get_local_properties(par, ipop)
    dg = vs = None
    if skewers:
        n, nr = par.nsources.get(ipop, 0), par.n_grid // 2
        dg, vs = np.zeros((n, nr), np.float32), np.zeros((n, nr), np.float32)
        check(par.lib.clr_srcs_get_skewers(par.ctx, C.c_int(ipop), _vp(dg), _vp(vs)))
    return srcs, dg, vs


def cstm_get_beam_properties(par: ParamCoLoRe, ipop: int, pos: np.ndarray) -> np.ndarray:
    """cstm.c:38-145 for pixels with unit vectors ``pos`` [npix,3] -> data[npix] (nadd = 1 everywhere)."""
    pos = np.ascontiguousarray(pos, np.float64)
    data = np.empty(pos.shape[0], np.float32)
    check(par.lib.clr_cstm_get_beam_properties(par.ctx, C.c_int(ipop), C.c_longlong(pos.shape[0]), _vp(pos), _vp(data)))
    return data


def lensing_get_beam_properties(par: ParamCoLoRe, r_sh, npp, pos: np.ndarray):
    """lensing.c:39-250 (fast-lensing shells). ``pos``: unit vectors of the finest shell [nbeams * npp[-1], 3]. Returns
    (data flat: shell ir = [nbeams][5 * npp[ir]], r_sh snapped to the radial sampling)."""
    r_sh = np.ascontiguousarray(r_sh, np.float32).copy()
    npp = np.ascontiguousarray(npp, np.int64)
    pos = np.ascontiguousarray(pos, np.float64)
    nbeams = pos.shape[0] // int(npp[-1])
    data = np.zeros(5 * nbeams * int(npp.sum()), np.float32)
    check(par.lib.clr_lensing_get_beam_properties(par.ctx, C.c_int(nbeams), C.c_int(len(r_sh)), _vp(r_sh), _vp(npp), _vp(pos),
                                                  _vp(data)))
    return data, r_sh


def srcs_lensing_from_shells(par: ParamCoLoRe, ipop: int, r_sh, nside_sh, node: int = 0, nnodes: int = 1):
    """srcs.c:666-723: lensing of the sources from the shells of the last lensing_get_beam_properties. Returns
    (srcs[n,9], number of sources outside the held base pixels)."""
    r_sh = np.ascontiguousarray(r_sh, np.float32)
    nside_sh = np.ascontiguousarray(nside_sh, np.int32)
    bad = C.c_longlong()
    check(par.lib.clr_srcs_lensing_from_shells(par.ctx, C.c_int(ipop), C.c_int(len(r_sh)), _vp(r_sh), _vp(nside_sh),
                                               C.c_int(node), C.c_int(nnodes), C.byref(bad)))
    return srcs_get_local_properties(par, ipop), int(bad.value)


def imap_set_cartesian(par: ParamCoLoRe, ipop: int = 0):
    """imap.c:135-245 -> (data[nr,npix], nadd[nr,npix]); the written map is data/nadd (io.c:738-741)."""
    nside, nr = par.imap_shells[ipop]
    npix = 12 * nside * nside
    data = np.empty((nr, npix), np.float32)
    nadd = np.empty((nr, npix), np.int32)
    check(par.lib.clr_imap_set_cartesian(par.ctx, C.c_int(ipop), _vp(data), _vp(nadd)))
    return data, nadd


def kappa_get_beam_properties(par: ParamCoLoRe, pos: np.ndarray, rf) -> np.ndarray:
    """kappa.c:39-175 for pixels with unit vectors ``pos`` [npix,3] and sorted plane radii ``rf``."""
    pos = np.ascontiguousarray(pos, np.float64)
    rf = np.ascontiguousarray(np.sort(np.asarray(rf, np.float32)), np.float32)
    data = np.empty((len(rf), pos.shape[0]), np.float32)
    check(par.lib.clr_kappa_get_beam_properties(par.ctx, C.c_longlong(pos.shape[0]), _vp(pos), C.c_int(len(rf)),
                                                _vp(rf), _vp(data)))
    return data


def isw_get_beam_properties(par: ParamCoLoRe, pos: np.ndarray, rf) -> np.ndarray:
    """isw.c:39-147."""
    pos = np.ascontiguousarray(pos, np.float64)
    rf = np.ascontiguousarray(np.sort(np.asarray(rf, np.float32)), np.float32)
    data = np.empty((len(rf), pos.shape[0]), np.float32)
    check(par.lib.clr_isw_get_beam_properties(par.ctx, C.c_longlong(pos.shape[0]), _vp(pos), C.c_int(len(rf)),
                                              _vp(rf), _vp(data)))
    return data


def write_catalog(par: ParamCoLoRe, ipop: int, fname: str, fmt: str = "ascii", n_threads: int = 0) -> float:
    """write_catalog (io.c:1019-1236), ASCII or FITS without lensing / skewers, from the device-resident records
    (clr_write_catalog: pinned chunked read-back + multi-threaded formatting). Returns the wall time in seconds."""
    sec = C.c_double()
    check(par.lib.clr_write_catalog(par.ctx, C.c_int(ipop), fname.encode(), C.c_int({"ascii": 0, "fits": 1}[fmt]),
                                    C.c_int(ipop), C.c_int(n_threads), C.byref(sec)))
    return sec.value


def write_healpix_map(fname: str, data, nside: int, nest: bool = False, nadd=None, listpix=None, n_threads: int = 0) -> float:
    """he_write_healpix_map (healpix_extra.c:4-57) + the shell loops of write_kappa / write_isw / write_imap
    (io.c:697-1017) through clr_write_healpix_map: ``data`` [num_pix] float32 (with ``listpix``: the local pixels of a
    shell, ``nadd``: their hit counts), NEST -> RING when ``nest``. Returns the wall time in seconds. No GPU needed."""
    from ._lib import load
    lib = load()
    data = np.ascontiguousarray(data, np.float32)
    nadd_c = None if nadd is None else np.ascontiguousarray(nadd, np.int32)
    list_c = None if listpix is None else np.ascontiguousarray(listpix, np.int64)      # `long *listpix` (common.h:211)
    sec = C.c_double()
    check(lib.clr_write_healpix_map(_vp(data), None if nadd_c is None else _vp(nadd_c), None if list_c is None else _vp(list_c),
                                    C.c_longlong(data.shape[0]), C.c_long(nside), C.c_int(int(nest)), fname.encode(),
                                    C.c_int(n_threads), C.byref(sec)))
    return sec.value
