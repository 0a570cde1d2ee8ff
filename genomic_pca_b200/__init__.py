"""genomic_pca_b200 -- B200-native hot path of genomic_pca behind a C ABI.

This Python package is a thin ctypes binding over ``libgpca.so`` (built from ``csrc/`` for
sm_100a; see ``include/gpca.h``) used by the tests and ``bench.py``.  It contains no
numerical fallback of any kind: if the library is missing, or there is no sm_100 GPU, every
compute call raises.
"""
from .binding import (GpcaError, Context, EigenSnpConfig, QcConfig, lib, library_path,  # noqa: F401
                      hwe_chi_squared_p_value, map_snps_to_ld_blocks)
from . import plink  # noqa: F401

__all__ = ["GpcaError", "Context", "EigenSnpConfig", "QcConfig", "lib", "library_path",
           "hwe_chi_squared_p_value", "map_snps_to_ld_blocks", "plink"]
