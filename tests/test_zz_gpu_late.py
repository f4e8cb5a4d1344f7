"""GPU parity tests added at the very end of round 2 (run once on a B200 on their own, 9 passed:
profiles/r2_late_gpu_tests.log -- the file sorts last because the full suite was not re-run with it).

1. Marker-level priors that the annotation update feeds to the sweep (BayesR.jl:28, MTBayesABC.jl:28-30,
   BayesABC.jl:17-23): `jwas_sweep_bayesr` / `jwas_sweep_mt1` with per_marker_pi against the oracle's contract
   sweep, bit-exact, and whole annotated chains through runMCMC on the B200 backend against the same host logic
   over the oracle backend.
2. BayesL! / BayesC0! (BayesC0L.jl:19-47) reference arithmetic (`jwo_bayesl_ref`) against the CUDA library run the
   way this backend runs them (BayesC step, pi = 0, marker variances sigma^2 * gamma_j): 1e-5 relative.
3. (added after that run, not yet executed on a B200) EBVs for genotyped individuals without phenotypes; multi-trait
   RR-BLUP chains; an uncentred chain.  Their Python side runs on the CPU through tests/fake_gpu_plugin.py."""
import numpy as np
import pytest

from helpers import Problem, uniform_starts
from test_gpu_sweep_parity import GAMMA, PI_R, jw  # noqa: F401  (jw: module fixture)

pytestmark = pytest.mark.gpu

ENGINES = [dict(engine=0, lag=0, chain_ctas=0), dict(engine=1, lag=2, chain_ctas=4)]


def _opts(g, o):
    for k, v in o.items():
        g.set_option(k, v)


@pytest.mark.parametrize("o", ENGINES)
def test_bayesr_marker_level_class_priors(jw, oracle, o):
    prob = Problem(oracle, 500, 1000, seed=131, missing=0.02)
    n, p = prob.n, prob.p
    starts = uniform_starts(p, 256)
    g = jw.GpuSweeper(prob.packed, n, 1)
    g.set_blocks(starts); _opts(g, o)
    yc, al, be, de = prob.fresh_state(); de[:] = 1
    g.put_ycorr(yc); g.put_state(al, be, de)
    rng = np.random.default_rng(5)
    snp_pi = rng.dirichlet([30.0, 2.0, 1.0, 0.5], size=p)
    vare = prob.vary * 0.5
    sigma = prob.vary * 0.5 / (prob.xpx.mean() / n * p * float(GAMMA @ PI_R))
    for it in range(1, 4):
        rc, _ = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, None, de,
                                      method=oracle.METHOD_R, nreps_mode=0, independent=False, vare=vare,
                                      sigmaSq=sigma, pi=snp_pi, gamma=GAMMA, seed=3, it=it, lag=o["lag"])
        assert rc == 0
        g.sweep_bayesr(jw.SCHED_EXACT, 1, vare, sigma, snp_pi, GAMMA, 3, it)
        ga, _, gd = g.get_state()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(g.get_ycorr().view(np.uint32), yc.view(np.uint32))
        snp_pi = rng.dirichlet([30.0, 2.0, 1.0, 0.5], size=p)      # the priors change every iteration
    assert (de > 1).sum() > 0
    g.close()


@pytest.mark.parametrize("o", ENGINES)
def test_two_trait_marker_level_joint_priors(jw, oracle, o):
    prob = Problem(oracle, 403, 600, seed=141, ntraits=2)
    n, p = prob.n, prob.p
    starts = uniform_starts(p, 200)
    g = jw.GpuSweeper(prob.packed, n, 2)
    g.set_blocks(starts); _opts(g, o)
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    R = np.array([[1.0, 0.3], [0.3, 1.2]]) * prob.vary * 0.5
    G = np.array([[1.0, 0.4], [0.4, 0.8]]) * prob.vary * 0.5 / (0.2 * prob.xpx.mean() / n * p)
    rng = np.random.default_rng(6)
    for it in range(1, 4):
        snp_pi = rng.dirichlet([7.0, 1.0, 1.0, 1.0], size=p)       # columns 00, 10, 01, 11
        rc, _ = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, be, de,
                                      method=oracle.METHOD_MT1, nreps_mode=0, independent=False, R=R, G=G,
                                      bigPi=snp_pi, seed=9, it=it, lag=o["lag"])
        assert rc == 0
        g.sweep_mt1(jw.SCHED_EXACT, R, G, snp_pi, 9, it)
        ga, gb, gd = g.get_state()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(gb.view(np.uint32), be.view(np.uint32))
        np.testing.assert_array_equal(g.get_ycorr().view(np.uint32), yc.view(np.uint32))
    assert de.sum() > 0
    g.close()


@pytest.mark.parametrize("case", ["BayesC", "BayesR", "BayesC2"])
def test_annotated_chain_matches_oracle_chain(case):
    import jwas_b200
    from oracle_backend import factory
    from test_annotations import _annotated_data
    from test_gpu_chain import assert_same
    two = case == "BayesC2"
    codes, ids, ph, A = _annotated_data(n=250, p=320, seed=51, ntraits=2 if two else 1)
    eqs = "y1 = intercept + geno\ny2 = intercept + geno" if two else "y1 = intercept + geno"
    Pi = {(0.0, 0.0): 0.45, (1.0, 0.0): 0.20, (0.0, 1.0): 0.15, (1.0, 1.0): 0.20} if two else (0.9 if case == "BayesC" else 0.0)
    outs = []
    for bf in (None, factory):
        geno = jwas_b200.get_genotypes(codes, False, method="BayesR" if case == "BayesR" else "BayesC", Pi=Pi,
                                       annotations=A, obsID=ids)
        model = jwas_b200.build_model(eqs, False, genotypes={"geno": geno})
        outs.append(jwas_b200.runMCMC(model, ph, chain_length=30, burnin=6, seed=77, _backend_factory=bf))
    assert_same(outs[0], outs[1])
    assert "annotation coefficients geno" in outs[0]


@pytest.mark.parametrize("lasso", [False, True])
def test_cuda_default_vs_reference_bayesl(jw, oracle, lasso):
    N, P = 500, 2000
    prob = Problem(oracle, N, P, seed=2026)
    rng = np.random.default_rng(21)
    sum2pq = float((prob.means.astype(np.float64) * (1 - prob.means / 2)).sum())
    v_res = float(np.float32(prob.vary / 2))
    v_eff = float(np.float32((prob.vary / 2) / sum2pq / (8 if lasso else 1)))     # MCMC_BayesianAlphabet.jl:70-74
    gamma = rng.gamma(1.0, 8.0, P) if lasso else np.array([1.0])
    ve = np.full(P, v_eff) * gamma
    g = jw.GpuSweeper(prob.packed, N, 1)
    _opts(g, dict(engine=1, lag=2, chain_ctas=4))
    g.set_blocks(uniform_starts(P, 256))
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    y_r = prob.ycorr0.copy(); a_r = np.zeros(P, np.float32)
    zr = np.random.default_rng(4)
    for it in range(1, 4):
        u, z = zr.random(P), zr.standard_normal(P)
        oracle.bayesl_ref(prob.X, prob.xpx, y_r, a_r, gamma, v_res, v_eff, z)
        g.sweep_bayesabc(jw.SCHED_EXACT, v_res, ve, np.zeros(P), 1, it, u, z)
        ga, _, gd = g.get_state()
        assert gd.sum() == P
        ra = np.abs(ga.astype(np.float64) - a_r).max() / np.abs(a_r).max()
        ry = np.abs(g.get_ycorr().astype(np.float64) - y_r).max() / np.abs(y_r).max()
        assert ra <= 1e-5 and ry <= 1e-5, (lasso, it, ra, ry)
    g.close()


def test_ebv_for_unphenotyped_individuals_matches_oracle_chain():
    """check_outputID / align_genotypes / getEBV (input_data_validation.jl:143-196, tools4genotypes.jl:288-296,
    output.jl:300-304): training on a phenotyped subset, EBVs for every genotyped individual through a second handle
    that holds their rows (not yet run on a B200; everything it calls -- jwas_set_marker_means, jwas_put_state,
    jwas_mul_alpha -- is covered by GPU-verified tests)."""
    import jwas_b200
    from oracle_backend import factory
    from test_api_chain import make_data
    from test_gpu_chain import assert_same
    codes, ids, ph = make_data(n=260, p=300, seed=29, missing=0.01)
    sub = ph.iloc[::-1].iloc[:200].reset_index(drop=True)
    outs = []
    for bf in (None, factory):
        geno = jwas_b200.get_genotypes(codes, 1.0, method="BayesC", Pi=0.9, obsID=ids)
        model = jwas_b200.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
        outs.append(jwas_b200.runMCMC(model, sub, chain_length=20, burnin=4, seed=77, _backend_factory=bf))
    assert list(outs[0]["EBV_y1"]["ID"]) == ids
    g, o = outs[0]["EBV_y1"], outs[1]["EBV_y1"]
    # the device product sums its rows in another order than the oracle's: equal to Float32 rounding
    np.testing.assert_allclose(g["EBV"].to_numpy(float), o["EBV"].to_numpy(float), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g["PEV"].to_numpy(float), o["PEV"].to_numpy(float), rtol=1e-3, atol=1e-6)
    for key in outs[0]:
        if not key.startswith("EBV_"):
            assert_same({key: outs[0][key]}, {key: outs[1][key]})


@pytest.mark.parametrize("constraint", [False, True])
def test_multitrait_rrblup_chain_matches_oracle_chain(constraint):
    """MTBayesC0! / megaBayesC0! through sampler I with the prior mass on the all-traits state / megaBayesABC with pi = 0
    (log prior of the other states = -inf on both sides).  Not yet run on a B200; the CPU bridge to `jwo_mtbayesl_ref`
    is tests/test_ref_bridge.py::test_contract_tracks_reference_multitrait_rrblup."""
    import jwas_b200
    from oracle_backend import factory
    from test_api_chain import make_data
    from test_gpu_chain import assert_same
    codes, ids, ph = make_data(n=220, p=260, seed=63, ntraits=2)
    G = np.array([[1.0, 0.0 if constraint else 0.4], [0.0 if constraint else 0.4, 1.0]])
    outs = []
    for bf in (None, factory):
        geno = jwas_b200.get_genotypes(codes, G, method="RR-BLUP", obsID=ids, constraint=constraint)
        model = jwas_b200.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.eye(2), genotypes={"geno": geno},
                                      constraint=constraint)
        outs.append(jwas_b200.runMCMC(model, ph, chain_length=16, burnin=4, seed=77, _backend_factory=bf))
    assert_same(outs[0], outs[1])
    assert (outs[0]["marker effects geno"]["Model_Frequency"] == 1.0).all()


def test_uncentred_chain_matches_oracle_chain():
    """center=false: the handle is given means of zero (jwas_set_marker_means) so that x_ij is the code itself.  Not yet
    run on a B200 (the external-means path itself is: test_gpu_sweep_parity.py, full-sample means)."""
    import jwas_b200
    from oracle_backend import factory
    from test_api_chain import make_data
    from test_gpu_chain import assert_same
    codes, ids, ph = make_data(n=240, p=280, seed=67)
    outs = []
    for bf in (None, factory):
        geno = jwas_b200.get_genotypes(codes, 1.0, method="BayesC", Pi=0.9, obsID=ids, center=False)
        model = jwas_b200.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
        outs.append(jwas_b200.runMCMC(model, ph, chain_length=20, burnin=4, seed=77, _backend_factory=bf))
    assert_same(outs[0], outs[1])
