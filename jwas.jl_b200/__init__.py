"""jwas_b200 -- B200-native backend for the JWAS.jl marker-effects Gibbs sweep.

Host-side mirror of the reference interface for this path (get_genotypes / build_model /
set_random / runMCMC with Genotypes / MME types), driving hand-written sm_100a CUDA kernels
through the C ABI in include/jwas_b200.h.  There is no CPU fallback.
"""
from ._lib import (GpuSweeper, JwasError, SweepStats, SCHED_EXACT, SCHED_BLOCK, SCHED_INDEPENDENT,
                   device_count, shard_range, SO_PATH)

__all__ = ["GpuSweeper", "JwasError", "SweepStats", "SCHED_EXACT", "SCHED_BLOCK", "SCHED_INDEPENDENT",
           "device_count", "shard_range", "SO_PATH"]
from .api import (get_genotypes, build_model, set_covariate, set_random, runMCMC, outputEBV, prepare_streaming_genotypes,
                  load_streaming_backend, Genotypes, MME, MCMCinfo, Variance, resolve_fast_blocks,
                  validate_fast_block_starts)
from . import mcmc
from .memory import estimate_marker_memory, check_marker_memory_guard, format_bytes_human
from .gwas import GWAS

__all__ += ["get_genotypes", "build_model", "set_covariate", "set_random", "runMCMC", "outputEBV", "prepare_streaming_genotypes",
            "load_streaming_backend", "Genotypes", "MME", "MCMCinfo", "Variance", "mcmc", "GWAS",
            "estimate_marker_memory", "check_marker_memory_guard", "format_bytes_human"]
