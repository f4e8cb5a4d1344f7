"""GPU parity against the REFERENCE arithmetic: the CUDA library in its shipped configuration (engine 1 =
persistent fused sweep kernel, lagged schedule, chain pipelined over two chain CTAs; look-ahead panels of 2048
and of 256 markers) side by side with the faithful `*_ref` restatements of the reference samplers
(oracle/jwas_oracle.c: jwo_bayesabc_ref, jwo_bayesabc_block_ref, jwo_bayesr_ref, jwo_bayesr_block_ref,
jwo_mtbayesabc_I_ref, jwo_mtbayesabc_II_ref, jwo_mtbayesabc_block_ref) at BASELINE.json configs[0] size
(500 x 2,000) on shared replayed draws.  Bar (north_star): inclusion indicators EQUAL after every sweep,
effects and ycorr within 1e-5 relative (tests/bridge.py defines the norms).  The bit-exact comparison with
the contract-arithmetic twin is tests/test_gpu_sweep_parity.py; this file is the claim about the reference."""
import numpy as np
import pytest

import bridge as B
from helpers import Problem

pytestmark = pytest.mark.gpu
N, P = 500, 2000
REL = 1e-5


@pytest.fixture(scope="module")
def jw():
    import jwas_b200
    assert jwas_b200.device_count() > 0, "no CUDA device: the gpu-marked tests need a B200"
    return jwas_b200


@pytest.fixture(scope="module")
def probs(oracle):
    return {1: Problem(oracle, N, P, seed=2026), 2: Problem(oracle, N, P, seed=2027, ntraits=2),
            "miss": Problem(oracle, N, P, seed=2028, missing=0.02)}


def run_case(jw, oracle, prob, method, schedule, panel, nsweeps, engine=1, lag=1, chain_ctas=2):
    t = prob.t
    hyp = B.Hyper(prob, method, 5)
    if schedule == "exact":
        ref_starts = np.array([0, P], dtype=np.int64)
        starts = np.array(list(range(0, P, panel)) + [P], dtype=np.int64)
    else:
        ref_starts = starts = B.fast_block_starts(N, P)
    g = jw.GpuSweeper(prob.packed, N, t)
    g.set_option("engine", engine); g.set_option("lag", lag); g.set_option("chain_ctas", chain_ctas)
    g.set_blocks(starts)
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    sr = B.ref_state(prob, method)
    rng = np.random.default_rng(3)
    worst = (0.0, 0.0)
    for it in range(1, nsweeps + 1):
        u, z = B.draws(rng, schedule, ref_starts, t, P)
        B.ref_sweep(oracle, prob, hyp, schedule, ref_starts, sr, u, z)
        B.gpu_sweep(jw, g, hyp, schedule, u, z, it)
        ga, gb, gd = g.get_state()
        eq, ra, ry = B.compare(sr, (g.get_ycorr(), ga, gb, gd), method)
        assert eq, f"{method}/{schedule}/panel {panel}: delta forks from the reference arithmetic at sweep {it}"
        assert ra <= REL and ry <= REL, (method, schedule, panel, it, ra, ry)
        worst = (max(worst[0], ra), max(worst[1], ry))
    assert np.count_nonzero(ga) > 5
    g.close()
    return worst


@pytest.mark.parametrize("panel", [2048, 256])
@pytest.mark.parametrize("method", B.METHODS)
def test_cuda_default_vs_reference_exact(jw, oracle, probs, method, panel):
    """BayesABC! / BayesR! / _MTBayesABC_samplerI! / _samplerII! (BayesABC.jl:60-80, BayesR.jl:45-97,
    MTBayesABC.jl:57-210) vs the shipped default."""
    prob = probs[2 if method.startswith("MT") else 1]
    run_case(jw, oracle, prob, method, "exact", panel, nsweeps=4)


@pytest.mark.parametrize("schedule", ["block", "independent"])
@pytest.mark.parametrize("method", B.METHODS)
def test_cuda_vs_reference_block_schedules(jw, oracle, probs, method, schedule):
    """BayesABC_block! / BayesR_block! / MTBayesABC_block! with nreps = block size and their independent-block
    forms (BayesABC.jl:118-255, BayesR.jl:111-273, MTBayesABC.jl:243-646); fast_blocks=true partition, b = 22."""
    prob = probs[2 if method.startswith("MT") else 1]
    run_case(jw, oracle, prob, method, schedule, 0, nsweeps=2)


@pytest.mark.parametrize("method", ["BayesC", "BayesR"])
def test_cuda_default_vs_reference_with_missing_calls(jw, oracle, probs, method):
    """code 3 = missing -> column mean -> 0 after centring (decode_marker!, streaming_genotypes.jl:978-1002)."""
    run_case(jw, oracle, probs["miss"], method, "exact", 256, nsweeps=3)


def test_cuda_engine0_vs_reference(jw, oracle, probs):
    run_case(jw, oracle, probs[1], "BayesC", "exact", 256, nsweeps=3, engine=0, lag=0, chain_ctas=0)
