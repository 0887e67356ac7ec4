"""ctypes binding of colore_b200/libcolore_b200.so (C ABI declared in include/colore_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C colore_b200/csrc``.
There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but
every compute call fails loudly when the shared library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# COLORE_B200_LIB: another build of the SAME library (kernel-variant experiments, tools/fft_bench.py)
SO_PATH = os.environ.get("COLORE_B200_LIB") or os.path.join(HERE, "libcolore_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "colore_b200.h")
NA = 5001

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)


class ClrParams(C.Structure):
    """struct clr_params of include/colore_b200.h (field order must match)."""
    _fields_ = [
        ("n_grid", C.c_int32), ("nz_here", C.c_int32), ("iz0_here", C.c_int32), ("dens_type", C.c_int32),
        ("bias_model", C.c_int32), ("do_smoothing", C.c_int32), ("smooth_potential", C.c_int32),
        ("nside_base", C.c_int32), ("numk", C.c_int32), ("seed_rng", C.c_uint32),
        ("l_box", C.c_float), ("reserved_", C.c_float),
        ("pos_obs", C.c_double * 3),
        ("r2_smooth", C.c_double), ("prefac_lensing", C.c_double),
        ("fgrowth_0", C.c_double), ("hubble_0", C.c_double), ("OmegaM", C.c_double), ("n_scal", C.c_double),
        ("r_max", C.c_double), ("glob_idr", C.c_double),
        ("logkmin", C.c_double), ("logkmax", C.c_double), ("idlogk", C.c_double),
        ("logkarr", c_double_p), ("pkarr", c_double_p),
        ("r_arr_r2z", c_double_p), ("z_arr_r2z", c_double_p), ("growth_d_arr", c_double_p),
        ("growth_d2_arr", c_double_p), ("growth_v_arr", c_double_p), ("growth_pd_arr", c_double_p),
        ("ihub_arr", c_double_p),
        ("a_arr_a2r", c_double_p), ("r_arr_a2r", c_double_p),
    ]


class ColoreError(RuntimeError):
    pass


_lib = None


def declared_symbols() -> list:
    """Every function declared in include/colore_b200.h."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clr_[a-z0-9_]+)\s*\(", src)))


def load() -> C.CDLL:
    """Load the CUDA shared library; raise if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ColoreError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the GPU path has no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    lib.clr_last_error.restype = C.c_char_p
    lib.clr_launch_count.restype = C.c_longlong
    lib.clr_create.argtypes = [C.POINTER(ClrParams), C.c_int, C.POINTER(C.c_void_p)]
    for name in declared_symbols():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        if name not in ("clr_last_error", "clr_launch_count", "clr_version", "clr_device_count"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise ColoreError(load().clr_last_error().decode())
