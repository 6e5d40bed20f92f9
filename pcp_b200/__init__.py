"""pcp_b200 -- B200-native propagation fixpoint behind libpcp's store surface.

The product is the C-ABI library `libpcp_b200.so` (include/pcp_b200.h); this package is
its Python binding plus workload generators.  Importing the package does not need a GPU;
creating an `Engine` does.
"""
from . import models
from ._capi import FALSE, TRUE, UNKNOWN, ContractViolation, PcpError
from .engine import ABI_SYMBOLS, LIB_PATH, Engine, consistency_batch, load_library, search_step_many

__all__ = ["Engine", "consistency_batch", "search_step_many", "models", "load_library", "ABI_SYMBOLS", "LIB_PATH", "PcpError", "ContractViolation",
           "TRUE", "FALSE", "UNKNOWN"]
