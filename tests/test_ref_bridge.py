"""Bridge between the two arithmetic flavours of the oracle, for EVERY method and schedule of the path:
the contract-arithmetic sweep (the bit-exact twin of the CUDA kernels) against the faithful `*_ref`
restatements of the reference samplers, at BASELINE.json configs[0] size (500 x 2,000) with shared replayed
draws (tests/bridge.py).  Bar = north_star's: inclusion indicators EQUAL after every sweep, effects and
ycorr within 1e-5 relative (measured: <= 6e-7; the floor is the reference's own Float32 sdot).

The shipped default schedule (lag = 1, look-ahead panels of 256 / 2048 markers) and the plain one (lag = 0)
are both bridged.  tests/test_gpu_ref_parity.py repeats this with the CUDA library as the second arm."""
import numpy as np
import pytest

import bridge as B
from helpers import Problem

N, P = 500, 2000
REL = 1e-5


@pytest.fixture(scope="module")
def probs(oracle):
    return {1: Problem(oracle, N, P, seed=2026), 2: Problem(oracle, N, P, seed=2027, ntraits=2)}


@pytest.mark.parametrize("schedule", B.SCHEDULES)
@pytest.mark.parametrize("method", B.METHODS)
def test_contract_tracks_reference_arithmetic(oracle, probs, method, schedule):
    t = 2 if method.startswith("MT") else 1
    prob = probs[t]
    hyp = B.Hyper(prob, method, 5)
    if schedule == "exact":
        ref_starts = np.array([0, P], dtype=np.int64)
        variants = [(np.array(list(range(0, P, 256)) + [P], dtype=np.int64), 1),     # shipped default: lagged panels
                    (np.array([0, P], dtype=np.int64), 1),                            # one panel of 2,000 (panel 2048)
                    (np.array(list(range(0, P, 256)) + [P], dtype=np.int64), 0)]
        nsweeps = 4
    else:
        ref_starts = B.fast_block_starts(N, P)             # fast_blocks=true: b = floor(sqrt(500)) = 22, nreps = b
        variants = [(ref_starts, 0)]
        nsweeps = 2
    for starts, lag in variants:
        sr = B.ref_state(prob, method)
        sc = prob.fresh_state()
        rng = np.random.default_rng(3)
        for it in range(1, nsweeps + 1):
            u, z = B.draws(rng, schedule, ref_starts, t, P)
            B.ref_sweep(oracle, prob, hyp, schedule, ref_starts, sr, u, z)
            B.contract_sweep(oracle, prob, hyp, schedule, starts, sc, u, z, it, lag=lag)
            eq, ra, ry = B.compare(sr, sc, method)
            assert eq, f"{method}/{schedule}: delta forks from the reference arithmetic at sweep {it} (lag={lag})"
            assert ra <= REL and ry <= REL, (method, schedule, it, lag, ra, ry)
        assert np.count_nonzero(sc[1]) > 5, "degenerate case: (almost) nothing in the model"


def test_block_refs_with_one_repetition_equal_dense_refs(oracle, probs):
    """BayesR_block! / MTBayesABC block samplers with nreps = 1 are the same chain as the dense samplers
    (block rhs corrected by Gram columns == dot against the updated ycorr; BayesR.jl:150,182): pins the new
    block restatements to the dense ones (Float32 rounding only)."""
    prob = probs[1]
    hyp = B.Hyper(prob, "BayesR", 5)
    starts = B.fast_block_starts(N, P)
    rng = np.random.default_rng(9)
    u, z = rng.random(P), rng.standard_normal(P)
    s1 = B.ref_state(prob, "BayesR"); s2 = B.ref_state(prob, "BayesR")
    oracle.bayesr_ref(prob.X, prob.xpx, s1[0], s1[1], s1[3], hyp.vare, hyp.sigma_sq, B.PI_R, B.GAMMA, u, z)
    oracle.bayesr_block_ref(prob.X, prob.xpx, starts, 1, False, s2[0], s2[1], s2[3], hyp.vare, hyp.sigma_sq, B.PI_R,
                            B.GAMMA, u, z)
    np.testing.assert_array_equal(s1[3], s2[3])
    np.testing.assert_allclose(s2[1], s1[1], atol=2e-6 * np.abs(s1[1]).max())
    prob = probs[2]
    for sampler, name in ((1, "MT1"), (2, "MT2")):
        hyp = B.Hyper(prob, name, 5)
        u, z = rng.random(2 * P), rng.standard_normal(2 * P)
        s1 = B.ref_state(prob, name); s2 = B.ref_state(prob, name)
        B.ref_sweep(oracle, prob, hyp, "exact", None, s1, u, z)
        oracle.mtbayesabc_block_ref(prob.X, prob.xpx, starts, 1, False, sampler, s2[0], s2[1], s2[2], s2[3], hyp.R, hyp.G,
                                    hyp.big_pi, u, z)
        np.testing.assert_array_equal(s1[3], s2[3])
        np.testing.assert_allclose(s2[1], s1[1], atol=2e-6 * np.abs(s1[1]).max())


@pytest.mark.parametrize("lasso", [False, True])
def test_contract_tracks_reference_bayesl(oracle, probs, lasso):
    """BayesL! / BayesC0! (BayesC0L.jl:19-47) against the way this backend runs them: the BayesC step with pi = 0 and
    marker variances sigma^2 * gamma_j (mcmc.run_chain).  Every marker is in the model, so there is no
    indicator to compare: effects and ycorr within 1e-5 relative after every sweep, default lagged panels and plain."""
    prob = probs[1]
    rng = np.random.default_rng(21)
    sum2pq = float((prob.means.astype(np.float64) * (1 - prob.means / 2)).sum())
    v_res = float(np.float32(prob.vary / 2))
    v_eff = float(np.float32((prob.vary / 2) / sum2pq / (8 if lasso else 1)))     # MCMC_BayesianAlphabet.jl:70-74
    gamma = rng.gamma(1.0, 8.0, P) if lasso else np.array([1.0])
    ve = np.full(P, v_eff) * gamma
    for starts, lag in ((np.array(list(range(0, P, 256)) + [P], dtype=np.int64), 2),
                        (np.array([0, P], dtype=np.int64), 0)):
        y_r = prob.ycorr0.copy(); a_r = np.zeros(P, np.float32)
        yc, al, be, de = prob.fresh_state()
        zr = np.random.default_rng(4)
        for it in range(1, 4):
            u, z = zr.random(P), zr.standard_normal(P)
            oracle.bayesl_ref(prob.X, prob.xpx, y_r, a_r, gamma, v_res, v_eff, z)
            rc, _ = oracle.sweep_contract(prob.packed, N, prob.means, prob.xpx, starts, yc, al, be, de,
                                          method=oracle.METHOD_ABC, nreps_mode=0, independent=False, vare=v_res,
                                          varEffects=ve, pi=np.zeros(P), seed=1, it=it, u=u, z=z, lag=lag)
            assert rc == 0
            assert de.sum() == P
            ra = np.abs(al.astype(np.float64) - a_r).max() / np.abs(a_r).max()
            ry = np.abs(yc.astype(np.float64) - y_r).max() / np.abs(y_r).max()
            assert ra <= REL and ry <= REL, (lasso, lag, it, ra, ry)


@pytest.mark.parametrize("method", ["BayesC", "BayesR", "MT1"])
def test_contract_tracks_reference_with_marker_level_priors(oracle, probs, method):
    """Annotated runs (BayesABC.jl:17-23 pi vector; BayesR.jl:28 and MTBayesABC.jl:28-30 snp_pi matrices): the
    contract sweep with marker-level priors against the `*_ref` samplers given the same priors, lag-2 panels."""
    t = 2 if method == "MT1" else 1
    prob = probs[t]
    hyp = B.Hyper(prob, method, 5)
    rng = np.random.default_rng(31)
    if method == "BayesC":
        hyp.pi = np.clip(rng.beta(30, 2, size=P), 0.5, 0.999)                 # pi_j
    elif method == "BayesR":
        prior = rng.dirichlet([60.0, 2.0, 1.0, 0.4], size=P)
    else:
        prior = rng.dirichlet([18.0, 1.0, 1.0, 1.0], size=P)                   # columns 00, 10, 01, 11
    starts = np.array(list(range(0, P, 256)) + [P], dtype=np.int64)
    sr = B.ref_state(prob, method)
    yc, al, be, de = prob.fresh_state()
    for it in range(1, 4):
        u, z = rng.random(t * P), rng.standard_normal(t * P)
        kw = dict(nreps_mode=0, independent=False, seed=1, it=it, u=u, z=z, lag=2)
        if method == "BayesC":
            B.ref_sweep(oracle, prob, hyp, "exact", None, sr, u, z)
            B.contract_sweep(oracle, prob, hyp, "exact", starts, (yc, al, be, de), u, z, it, lag=2)
        elif method == "BayesR":
            oracle.bayesr_ref(prob.X, prob.xpx, sr[0], sr[1], sr[3], hyp.vare, hyp.sigma_sq, prior, B.GAMMA, u, z)
            rc, _ = oracle.sweep_contract(prob.packed, N, prob.means, prob.xpx, starts, yc, al, be, de,
                                          method=oracle.METHOD_R, vare=hyp.vare, sigmaSq=hyp.sigma_sq, pi=prior,
                                          gamma=B.GAMMA, **kw)
            assert rc == 0
        else:
            oracle.mtbayesabc_I_ref(prob.X, prob.xpx, sr[0], sr[1], sr[2], sr[3], hyp.R, hyp.G, prior, u, z)
            rc, _ = oracle.sweep_contract(prob.packed, N, prob.means, prob.xpx, starts, yc, al, be, de,
                                          method=oracle.METHOD_MT1, R=hyp.R, G=hyp.G, bigPi=prior, **kw)
            assert rc == 0
        eq, ra, ry = B.compare(sr, (yc, al, be, de), method)
        assert eq, f"{method}: delta forks from the reference arithmetic at sweep {it}"
        assert ra <= REL and ry <= REL, (method, it, ra, ry)
    assert np.count_nonzero(al) > 5


def test_contract_tracks_reference_multitrait_rrblup(oracle, probs):
    """MTBayesC0! = MTBayesL! with gamma = [1.0] (MTBayesC0L.jl:6-58): every marker in the model for every trait.  This
    backend runs it as sampler I with all the prior mass on the all-traits state (constraint=false) and as
    megaBayesABC! with pi = 0 per trait (constraint=true: megaBayesC0!, BayesC0L.jl:13-17 = single-trait BayesL! per
    trait).  Effects and ycorr within 1e-5 relative of the reference arithmetic."""
    prob = probs[2]
    hyp = B.Hyper(prob, "MT1", 5)
    G = hyp.G * 0.05                        # all 2,000 markers carry the variance: a smaller per-marker share
    big = np.array([0.0, 0.0, 0.0, 1.0])
    starts = np.array(list(range(0, P, 256)) + [P], dtype=np.int64)
    y_r = prob.ycorr0.copy(); a_r = np.zeros((2, P), np.float32)
    yc, al, be, de = prob.fresh_state(); de[:] = 1
    zr = np.random.default_rng(4)
    for it in range(1, 4):
        u, z = zr.random(2 * P), zr.standard_normal(2 * P)
        oracle.mtbayesl_ref(prob.X, prob.xpx, y_r, a_r, [1.0], hyp.R, G, z)
        rc, _ = oracle.sweep_contract(prob.packed, N, prob.means, prob.xpx, starts, yc, al, be, de, method=oracle.METHOD_MT1,
                                      nreps_mode=0, independent=False, R=hyp.R, G=G, bigPi=big, seed=1, it=it, u=u, z=z, lag=2)
        assert rc == 0 and de.sum() == 2 * P
        ra = np.abs(al.astype(np.float64) - a_r.reshape(-1)).max() / np.abs(a_r).max()
        ry = np.abs(yc.astype(np.float64) - y_r).max() / np.abs(y_r).max()
        assert ra <= REL and ry <= REL, (it, ra, ry)
    # constraint=true: one single-trait BayesL! per trait with the diagonal variances
    vare = np.diag(hyp.R).copy(); ve = np.diag(G).copy()
    y_r = prob.ycorr0.copy(); a_r = np.zeros((2, P), np.float32)
    yc, al, be, de = prob.fresh_state(); de[:] = 1
    for it in range(1, 3):
        u, z = zr.random(2 * P), zr.standard_normal(2 * P)
        for k in range(2):
            oracle.bayesl_ref(prob.X, prob.xpx, y_r[k * N:(k + 1) * N], a_r[k], [1.0], float(np.float32(vare[k])),
                              float(np.float32(ve[k])), z[k * P:(k + 1) * P])
        rc, _ = oracle.sweep_contract(prob.packed, N, prob.means, prob.xpx, starts, yc, al, be, de, method=oracle.METHOD_MEGA,
                                      nreps_mode=0, independent=False, R=np.diag(vare), G=np.diag(ve), bigPi=np.zeros(2),
                                      seed=1, it=it, u=u, z=z, lag=2)
        assert rc == 0 and de.sum() == 2 * P
        ra = np.abs(al.astype(np.float64) - a_r.reshape(-1)).max() / np.abs(a_r).max()
        ry = np.abs(yc.astype(np.float64) - y_r).max() / np.abs(y_r).max()
        assert ra <= REL and ry <= REL, ("mega", it, ra, ry)
