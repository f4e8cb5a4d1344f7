"""Marker-memory estimator and guard: the reference's own known answers (test/unit/test_memory_guardrails.jl:8-60,
77-117, 120-160 -- SURVEY.md §8c pin 8) for the modes it has, the :gpu mode this backend adds, and the precheck in
runMCMC (JWAS.jl:415-458)."""
import numpy as np
import pytest

import jwas_b200 as jw
from oracle_backend import factory
from test_api_chain import make_data


def test_estimator_reference_pins():
    est = jw.estimate_marker_memory(10, 20, element_bytes=4, has_nonunit_weights=False, block_starts=False)
    assert est["bytes_X"] == 10 * 20 * 4 and est["bytes_xRinvArray"] == 0 and est["bytes_XRinvArray"] == 0
    assert est["bytes_XpRinvX"] == 0 and est["bytes_xpRinvx"] == 20 * 4
    assert est["bytes_total"] == est["bytes_X"] + est["bytes_xpRinvx"]
    est = jw.estimate_marker_memory(10, 20, element_bytes=8, has_nonunit_weights=True, block_starts=False)
    assert est["bytes_X"] == 10 * 20 * 8 and est["bytes_xRinvArray"] == 10 * 20 * 8 and est["bytes_XRinvArray"] == 0
    assert est["bytes_XpRinvX"] == 0 and est["bytes_xpRinvx"] == 20 * 8
    est = jw.estimate_marker_memory(10, 20, element_bytes=4, has_nonunit_weights=False, block_starts=[1, 6, 11, 16])
    assert est["bytes_XRinvArray"] == 0 and est["bytes_XpRinvX"] == 100 * 4          # blocks of 5: sum(s_i^2) = 100
    dense = jw.estimate_marker_memory(100, 200, element_bytes=4, storage_mode="dense")
    stream = jw.estimate_marker_memory(100, 200, element_bytes=4, storage_mode="stream")
    assert stream["bytes_X"] == 0 and stream["bytes_decode_buffer"] == 100 * 4 and stream["bytes_marker_means"] == 200 * 4
    assert stream["bytes_xpRinvx"] == 200 * 4 and stream["bytes_packed_row_buffer"] == 25
    assert stream["bytes_total"] < dense["bytes_total"]
    # products beyond 64 bits do not overflow
    assert jw.estimate_marker_memory(2 ** 40, 2 ** 30, element_bytes=8)["bytes_X"] == 2 ** 73
    for bad in (dict(nObs=-1, nMarkers=1, element_bytes=4), dict(nObs=1, nMarkers=1, element_bytes=0),
                dict(nObs=1, nMarkers=1, element_bytes=4, storage_mode="tape")):
        with pytest.raises(jw.JwasError):
            jw.estimate_marker_memory(bad.pop("nObs"), bad.pop("nMarkers"), **bad)


def test_gpu_mode_matches_the_layout_in_design_md():
    """cfg2 (50,000 x 600,000, panels of 4096, lag 2): 7.5 GB packed + 7.5 GB tiled, Gram + two cross-Gram sets of
    9.8 GB each; at 8 GPUs every rank stores an eighth of the rows of cfg5."""
    starts = list(range(1, 600001, 4096))
    est = jw.estimate_marker_memory(50000, 600000, element_bytes=4, block_starts=starts, storage_mode="gpu", lag=2)
    assert est["bytes_packed"] == 600000 * 12512 == est["bytes_tiled"]               # cld(50,000, 4) = 12,500 -> pitch 12,512
    assert abs(est["bytes_XpRinvX"] / 1e9 - 9.8) < 0.1 and abs(est["bytes_cross_gram"] / 1e9 - 19.6) < 0.3
    assert 40e9 < est["bytes_total"] < 50e9
    one = jw.estimate_marker_memory(400000, 1000000, element_bytes=4, block_starts=list(range(1, 1000001, 4096)),
                                    storage_mode="gpu", lag=2)
    eight = jw.estimate_marker_memory(400000, 1000000, element_bytes=4, block_starts=list(range(1, 1000001, 4096)),
                                      storage_mode="gpu", lag=2, world=8)
    assert one["bytes_packed"] == 1000000 * 100000 and eight["bytes_packed"] == 1000000 * 12512
    assert one["bytes_total"] > 180e9 > eight["bytes_total"]      # cfg5 does not fit one B200 with its tiled copy, fits at 8
    assert eight["bytes_XpRinvX"] == one["bytes_XpRinvX"]         # marker-indexed data is replicated


def test_guard_modes_and_arguments():
    kw = dict(ratio=0.5, estimated_bytes=600, total_memory_bytes=1000, context_string="test")
    with pytest.raises(jw.JwasError, match="exceeds configured guard threshold"):
        jw.check_marker_memory_guard(mode="error", **kw)
    with pytest.warns(UserWarning, match="exceeds configured guard threshold"):
        assert jw.check_marker_memory_guard(mode="warn", **kw) == "warned"
    assert jw.check_marker_memory_guard(mode="off", **kw) == "skipped"
    assert jw.check_marker_memory_guard(mode=":error", ratio=0.5, estimated_bytes=400, total_memory_bytes=1000) == "ok"
    with pytest.raises(jw.JwasError, match="must be one of"):
        jw.check_marker_memory_guard(mode="unknown", ratio=0.5, estimated_bytes=1, total_memory_bytes=1000)
    with pytest.raises(jw.JwasError, match="memory_guard_ratio"):
        jw.check_marker_memory_guard(mode="error", ratio=0.0, estimated_bytes=1, total_memory_bytes=1000)
    assert jw.format_bytes_human(0) == "0.00 B" and jw.format_bytes_human(1536) == "1.50 KiB"
    assert jw.format_bytes_human(3 * 1024 ** 3) == "3.00 GiB"


def test_runmcmc_precheck():
    """test_memory_guardrails.jl:120-160: a tiny ratio stops the run early in :error mode, :off and :warn proceed."""
    codes, ids, ph = make_data(n=60, p=40, seed=3)

    def mk():
        geno = jw.get_genotypes(codes, 1.0, method="BayesC", obsID=ids)
        return jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    with pytest.raises(jw.JwasError, match="storage=:gpu, nObs=60, nMarkers=40"):
        jw.runMCMC(mk(), ph, chain_length=10, seed=123, memory_guard="error", memory_guard_ratio=1e-12, _backend_factory=factory)
    out = jw.runMCMC(mk(), ph, chain_length=10, seed=123, memory_guard="off", memory_guard_ratio=1e-12, _backend_factory=factory)
    assert len(out["EBV_y1"]) == 60
    with pytest.warns(UserWarning, match="guard threshold"):
        jw.runMCMC(mk(), ph, chain_length=10, seed=123, memory_guard="warn", memory_guard_ratio=1e-12, _backend_factory=factory)
