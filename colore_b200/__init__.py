"""colore_b200 -- B200-native (sm_100a CUDA) density-field -> catalogue/maps path of CoLoRe.

The package holds only what that path needs: ``csrc/`` (CUDA kernels + the C ABI of
include/colore_b200.h, built into ``libcolore_b200.so``), the ctypes binding, the host-side mirror
of the reference's run flow (``pipeline``), the slab decomposition (``dist``), host table
construction (``cosmo``) and synthetic inputs (``inputs``). There is no CPU fallback.
"""
from . import cosmo, dist, healpix, inputs, predictions  # noqa: F401
from ._lib import ColoreError, declared_symbols, load  # noqa: F401
from .pipeline import *  # noqa: F401,F403
from .pipeline import ParamCoLoRe  # noqa: F401

__version__ = "0.1.0"
