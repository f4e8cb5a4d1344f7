"""Marker-path memory estimate and guard (markers/tools4genotypes.jl:90-236, JWAS.jl:415-458).

`estimate_marker_memory` mirrors the reference for its own storage modes (:dense, :stream) and adds :gpu, the HBM
footprint of a handle of this backend: the packed 2-bit image and its tiled streaming copy, the Gram blocks
of the partition plus one cross-Gram set per lag, per-marker statistics and chain constants, sampler state and
accumulators, the exact int64 partial rhs, commit records, ycorr and its fixed-point image.  With several GPUs a rank
stores 1 / world of the rows; everything indexed by marker is replicated."""
import math
import os
import warnings

from ._lib import JwasError


def _block_sizes(block_starts, n_markers):
    """1-based block starts as the reference passes them (JWAS.jl:293-316) -> block sizes."""
    starts = [int(b) for b in block_starts]
    return [max(0, (n_markers if i + 1 == len(starts) else starts[i + 1] - 1) - s + 1) for i, s in enumerate(starts)]


def estimate_marker_memory(nObs, nMarkers, *, element_bytes, has_nonunit_weights=False, block_starts=False,
                           storage_mode="dense", n_traits=1, lag=0, world=1):
    if nObs < 0 or nMarkers < 0:
        raise JwasError("nObs and nMarkers must be non-negative.")
    if element_bytes <= 0:
        raise JwasError("element_bytes must be a positive integer.")
    if storage_mode not in ("dense", "stream", "gpu"):
        raise JwasError("storage_mode must be :dense, :stream or :gpu.")
    est = dict(bytes_X=0, bytes_xRinvArray=0, bytes_XRinvArray=0, bytes_XpRinvX=0, bytes_xpRinvx=0, bytes_decode_buffer=0,
               bytes_packed_row_buffer=0, bytes_marker_means=0)
    sizes = _block_sizes(block_starts, nMarkers) if block_starts is not False else []
    if storage_mode == "dense":
        est["bytes_X"] = nObs * nMarkers * element_bytes
        est["bytes_xpRinvx"] = nMarkers * element_bytes
        est["bytes_xRinvArray"] = est["bytes_X"] if has_nonunit_weights else 0
        # XRinvArray is not persisted in block mode; XpRinvX holds the per-block Gram matrices: sum(s_i^2)
        est["bytes_XpRinvX"] = sum(s * s for s in sizes) * element_bytes
    elif storage_mode == "stream":                      # O(N + P) working memory
        est["bytes_decode_buffer"] = nObs * element_bytes
        est["bytes_packed_row_buffer"] = -(-nObs // 4)
        est["bytes_marker_means"] = nMarkers * element_bytes
        est["bytes_xpRinvx"] = nMarkers * element_bytes
    else:
        t, p = int(n_traits), int(nMarkers)
        n_local = -(-int(nObs) // max(1, int(world)))
        pitch = -(-(-(-n_local // 4)) // 16) * 16           # column pitch: cld(rows, 4) bytes padded to 16
        gram = sum(s * s for s in sizes) * 4                  # Float32 Gram blocks of the partition (exact schedule: panels)
        cross = sum(a * b for a, b in zip(sizes[1:], sizes[:-1])) * 4
        est.update(
            bytes_packed=p * pitch, bytes_tiled=p * pitch,
            bytes_XpRinvX=gram, bytes_cross_gram=int(lag) * cross,
            bytes_marker_means=p * 4, bytes_xpRinvx=p * 4,
            bytes_marker_stats=p * (4 + 4 + 3 * 4),            # column sums, observed counts, call counts
            bytes_state=3 * t * p * 4, bytes_accumulators=3 * t * p * 4,
            bytes_partial_rhs=2 * t * p * 8, bytes_chain_constants=6 * p * 8 + p * 4 + 2 * t * p * 8,
            bytes_records=(p + len(sizes) * 4 + 64) * 8,
            bytes_ycorr=2 * t * int(nObs) * 4)
    est["bytes_total"] = sum(v for k, v in est.items() if k != "bytes_total")
    return est


def format_bytes_human(nbytes):
    if nbytes < 0:
        raise JwasError("bytes must be non-negative.")
    units = ("B", "KiB", "MiB", "GiB", "TiB", "PiB", "EiB")
    value, i = float(nbytes), 0
    while value >= 1024 and i < len(units) - 1:
        value /= 1024
        i += 1
    return "%.2f %s" % (value, units[i])


def check_marker_memory_guard(*, mode="error", ratio=0.8, estimated_bytes, total_memory_bytes, context_string=""):
    """tools4genotypes.jl:197-236 -> "ok" | "warned" | "skipped", or raises."""
    mode = str(mode).lstrip(":")
    if mode not in ("error", "warn", "off"):
        raise JwasError("memory_guard must be one of :error, :warn, or :off.")
    if not 0 < ratio <= 1:
        raise JwasError("memory_guard_ratio must satisfy 0 < memory_guard_ratio <= 1.")
    if estimated_bytes < 0 or total_memory_bytes <= 0:
        raise JwasError("estimated_bytes must be >= 0 and total_memory_bytes must be > 0.")
    if mode == "off":
        return "skipped"
    threshold = int(math.floor(float(total_memory_bytes) * float(ratio)))
    if estimated_bytes <= threshold:
        return "ok"
    msg = ("Estimated marker memory usage exceeds configured guard threshold.\n"
           f"context: {context_string}\n"
           f"estimated: {format_bytes_human(estimated_bytes)}\n"
           f"threshold ({ratio * 100}% of memory): {format_bytes_human(threshold)}\n"
           f"device / system memory: {format_bytes_human(total_memory_bytes)}\n"
           "Set memory_guard=:warn or :off to override, or reduce model/data size.")
    if mode == "warn":
        warnings.warn(msg)
        return "warned"
    raise JwasError(msg)


B200_HBM_BYTES = 180 * 1000 ** 3      # the capacity the guard assumes for a B200 unless JWAS_B200_HBM_BYTES says otherwise


def device_memory_bytes():
    return int(os.environ.get("JWAS_B200_HBM_BYTES", B200_HBM_BYTES))
