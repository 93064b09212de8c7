"""pedigreesim_b200 — B200 (sm_100a) meiosis + genotype emission for PedigreeSim's hot path.

The product is the C-ABI CUDA library libpsim.so (include/psim.h); this package is its thin
Python binding plus the sharding helpers the benchmark uses. No CPU fallback exists.
"""
from ._lib import (DOSE_MAX_ALLELE, DOSE_MIN_ALLELE, EMIT_ASYNC, EMIT_PACKED4, MEM_DEVICE, MEM_HOST, PATH_AUTO, PATH_GENERIC,
                   PATH_ROWS, PATH_TILES, PSIM_ECAPACITY, PSIM_ECUDA, PSIM_EINVAL, PSIM_ESTATE, EmitTargets, PsimContext,
                   PsimError, TextTargets, emit_path, load, PARTITION_BLOCK, PARTITION_LINEAGE, ShardGroup, shard_plan,
                   shard_unique_id)

__all__ = ["PsimContext", "PsimError", "EmitTargets", "load", "DOSE_MAX_ALLELE", "DOSE_MIN_ALLELE", "MEM_HOST",
           "MEM_DEVICE", "EMIT_ASYNC", "EMIT_PACKED4", "PATH_AUTO", "PATH_ROWS", "PATH_TILES", "PATH_GENERIC", "emit_path",
           "PARTITION_BLOCK", "PARTITION_LINEAGE", "ShardGroup", "shard_plan", "shard_unique_id"]
