"""corticall_b200 -- B200 (sm_100a) implementation of Corticall's data-parallel k-mer hot path.

The product is `libcorticall_cuda.so` (C ABI in include/corticall_cuda.h; CUDA sources in csrc/).  `host/`
mirrors the reference's Java class-library surface for that path (CortexGraph, CortexRecord, CanonicalKmer,
FindROIs, the Call helpers) on top of the C ABI.  There is no CPU fallback.
"""
from ._native import (CC_ALGO_AUTO, CC_ALGO_BSEARCH, CC_ALGO_MERGE, CortexJDKException, device_count, launch_count, lib,
                      set_option)
from .host.commands import (CallHelpers, CortexCollection, CortexVertex, CovStats, FindLowCoverage, FindROIs, FindShared, Join,
                            RecoverExcludedKmers, Remove, Sort)
from .host.cortex import CortexColor, CortexGraph, CortexHeader, CortexMap, CortexRecord, ShardedCortexGraph, packCanonical
from .host.kmer import CanonicalKmer, CortexBinaryKmer, CortexByteKmer, SequenceUtils

__all__ = ["CortexGraph", "ShardedCortexGraph", "CortexMap", "CortexRecord", "CortexHeader", "CortexColor", "CanonicalKmer", "CortexByteKmer",
           "CortexBinaryKmer", "SequenceUtils", "FindROIs", "Join", "Remove", "Sort", "FindLowCoverage", "FindShared", "RecoverExcludedKmers", "CovStats", "CortexCollection", "CallHelpers", "CortexVertex", "CortexJDKException", "packCanonical", "lib", "launch_count",
           "set_option", "device_count", "CC_ALGO_AUTO", "CC_ALGO_BSEARCH", "CC_ALGO_MERGE"]
