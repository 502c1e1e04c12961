"""tiktak_b200 -- B200-native implementation of TikTak's Sobol pre-test + batched local-search hot path.

Compute lives in libtiktak_cuda.so (hand-written sm_100a CUDA behind the C ABI of
include/tiktak_cuda.h); this package is the host mirror of the reference's stage procedures.
"""
from ._abi import (ALG_AMOEBA, ALG_DFNLS, OBJ_GRIEWANK, OBJ_RASTRIGIN, OBJ_ROSENBROCK, OBJ_SMM, OBJ_TEMPLATE,
                   SOBOL_BLOCK)
from .config import TikTakConfig, baseline_config, make_config, parse_config, smm_config
from .lib import TikTakError
from .search import TikTakSearch

__all__ = ["TikTakSearch", "TikTakConfig", "TikTakError", "parse_config", "make_config", "baseline_config", "smm_config",
           "OBJ_TEMPLATE", "OBJ_GRIEWANK", "OBJ_RASTRIGIN", "OBJ_ROSENBROCK", "OBJ_SMM", "ALG_AMOEBA", "ALG_DFNLS",
           "SOBOL_BLOCK"]
