"""pedigreesim_b200 — B200 (sm_100a) meiosis + genotype emission for PedigreeSim's hot path.

The product is the C-ABI CUDA library libpsim.so (include/psim.h); this package is its thin
Python binding plus the sharding helpers the benchmark uses. No CPU fallback exists.
"""
from ._lib import (DOSE_MAX_ALLELE, DOSE_MIN_ALLELE, MEM_DEVICE, MEM_HOST, PSIM_ECAPACITY, PSIM_ECUDA, PSIM_EINVAL, PSIM_ESTATE,
                   EmitTargets, PsimContext, PsimError, TextTargets, load)

__all__ = ["PsimContext", "PsimError", "EmitTargets", "load", "DOSE_MAX_ALLELE", "DOSE_MIN_ALLELE", "MEM_HOST",
           "MEM_DEVICE"]
