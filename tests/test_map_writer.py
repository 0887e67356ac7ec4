"""CPU suite: the native HEALPix map writer (clr_write_healpix_map, colore_b200/csrc/clr_io.cu) against
  (1) the unmodified reference's he_write_healpix_map (healpix_extra.c:4-57, from oracle/_ref/libcolore_ref.so, whose
      FITS / HEALPix calls go to the third-party stand-ins): same bytes, NEST and RING input, several resolutions;
  (2) an independent statement of the RING order: along the file the colatitude never decreases and the longitude
      increases inside a ring (unit vectors from colore_b200.healpix.pix2vec_nest, itself pinned by the golden shells);
  (3) the shell loops of write_kappa / write_imap (io.c:697-1017) restated in numpy.
No GPU involved: the maps are host arrays at the boundary."""
import ctypes as C
import os

import numpy as np
import pytest

import colore_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcolore_ref.so")


def _table(raw):
    """(header cards of the table HDU, float32 column) of a one-column BINTABLE file"""
    assert len(raw) % 2880 == 0
    cards = [raw[2880 + i:2880 + i + 80].decode() for i in range(0, 2880, 80)]
    keys = {c[:8].strip(): c[10:].split("/")[0].strip() for c in cards if c[8:10] == "= "}
    n = int(keys["NAXIS2"])
    col = np.frombuffer(raw[5760:5760 + 4 * n], dtype=">f4").astype(np.float32)
    assert raw[5760 + 4 * n:] == b"\0" * (len(raw) - 5760 - 4 * n)
    return cards, keys, col


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libcolore_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("nside", [1, 2, 16, 128])
@pytest.mark.parametrize("nest", [0, 1])
def test_same_bytes_as_the_reference_writer(tmp_path, nside, nest):
    ref = C.CDLL(REF_SO)
    npix = 12 * nside * nside
    m = np.random.default_rng(nside + nest).normal(size=npix).astype(np.float32)
    f_ref, f_new = str(tmp_path / "ref.fits"), str(tmp_path / "new.fits")
    buf = m.copy()                                               # he_nest2ring_inplace reorders its argument
    ptr = (C.POINTER(C.c_float) * 1)(buf.ctypes.data_as(C.POINTER(C.c_float)))
    ref.he_write_healpix_map.argtypes = [C.POINTER(C.POINTER(C.c_float)), C.c_int, C.c_long, C.c_char_p, C.c_int]
    ref.he_write_healpix_map(ptr, 1, nside, ("!" + f_ref).encode(), nest)
    for nt in (1, 3):
        cb.write_healpix_map("!" + f_new, m, nside, nest=bool(nest), n_threads=nt)
        assert open(f_new, "rb").read() == open(f_ref, "rb").read()


@pytest.mark.parametrize("nside", [1, 4, 64])
def test_ring_order_of_the_file(tmp_path, nside):
    """NEST input whose value IS its NEST index: the file then lists ring2nest(p) for p = 0, 1, ... -- a permutation
    whose unit vectors walk down the sphere ring by ring, eastwards inside every ring (4 i pixels in cap ring i,
    4 nside in the belt)."""
    npix = 12 * nside * nside
    f = str(tmp_path / "m.fits")
    cb.write_healpix_map(f, np.arange(npix, dtype=np.float32), nside, nest=True)
    _, keys, col = _table(open(f, "rb").read())
    assert keys["ORDERING"] == "'RING'" or keys["ORDERING"].strip("' ") == "RING"
    assert int(keys["NSIDE"]) == nside
    perm = col.astype(np.int64)
    assert np.array_equal(np.sort(perm), np.arange(npix))
    v = cb.healpix.pix2vec_nest(nside, perm)
    z, phi = v[:, 2], np.mod(np.arctan2(v[:, 1], v[:, 0]), 2 * np.pi)
    assert np.all(np.diff(z) <= 1e-12)
    start = 0
    for ring in range(1, 4 * nside):
        n_in_ring = 4 * min(ring, nside, 4 * nside - ring)
        zz, pp = z[start:start + n_in_ring], phi[start:start + n_in_ring]
        assert np.ptp(zz) < 1e-12 and np.all(np.diff(pp) > 0)
        start += n_in_ring
    assert start == npix


def test_shell_loops_of_write_kappa(tmp_path):
    """io.c:820-845: local pixels scattered through listpix, divided by their hit counts where there are hits"""
    nside = 8
    npix = 12 * nside * nside
    rng = np.random.default_rng(5)
    listpix = rng.permutation(npix)[: npix - 50].astype(np.int64)          # 50 pixels nobody owns
    data = rng.normal(size=listpix.size).astype(np.float32)
    nadd = rng.integers(0, 4, size=listpix.size).astype(np.int32)
    f = str(tmp_path / "k.fits")
    cb.write_healpix_map(f, data, nside, nest=False, nadd=nadd, listpix=listpix)
    _, _, col = _table(open(f, "rb").read())
    want = np.zeros(npix, np.float32)
    hits = np.zeros(npix, np.int32)
    want[listpix] = data
    hits[listpix] = nadd
    want[hits > 0] /= hits[hits > 0].astype(np.float32)
    assert np.array_equal(col, want)


def test_bad_arguments():
    with pytest.raises(cb.ColoreError):
        cb.write_healpix_map("/tmp/x.fits", np.zeros(10, np.float32), 3)                 # nside not a power of two
    with pytest.raises(cb.ColoreError):
        cb.write_healpix_map("/tmp/x.fits", np.zeros(10, np.float32), 2)                 # 10 pixels, no pixel list
    with pytest.raises(cb.ColoreError, match="open file"):
        cb.write_healpix_map("/nonexistent_dir/x.fits", np.zeros(48, np.float32), 2)
